// chz_oracle — CPU restatement (double precision, OpenMP) of the reference hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (sdr_channelizer_b200/, libchannelizer.so, the
// CLI) may import, link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker and the reported CPU baseline.
//
// What it follows (paths relative to the reference root, cwozny/sdr_channelizer):
//   reader      matlab/convert_my_iq_to_mat.m:38-102, header struct cpp/IqPacket.h:9-25
//   normalise   matlab/create_pdws_channelized.m:35-38
//   trim/shift  matlab/create_pdws_channelized.m:52-62
//   PDWs        matlab/create_pdws_channelized.m:67-136
//
// PARITY STATUS
//   * Header layout (R1) is PINNED: tests check this parser against a file written by the
//     reference's own IqPacket struct, compiled from /root/reference/cpp/IqPacket.h into
//     oracle/_ref/ (see oracle/Makefile, oracle/ref_iqpacket_writer.cpp) and committed as
//     tests/golden/ref_iqpacket_fmt3.iq.
//   * Channelizer arithmetic is **parity unpinned**: the reference calls MathWorks' closed-source
//     DSP System Toolbox object dsp.Channelizer (create_pdws_channelized.m:33,57;
//     channelizer_example.m:31,56), which is not in the reference tree, has no pinned version and
//     cannot run here (no MATLAB/Octave).  The restatement below is the textbook polyphase
//     analysis bank with dsp.Channelizer's documented defaults (12 taps/band, 80 dB Kaiser
//     prototype); it is validated against a direct-form evaluation of its defining sum and against
//     numpy/scipy, not against MATLAB output.  The reference ships no golden vectors or tests.
//   * PDW extraction is a line-for-line restatement of the script; the reference holds no vectors
//     for it either, so its known-answer tests are hand-built traces (tests/test_oracle.py).
//
// Channelizer definition:
//   y_k[m] = e^{-j 2 pi k m D / M} * sum_{n=0}^{L-1} h[n] e^{+j 2 pi k n / M} x[m D - n],  x[n<0] = 0
//   rows m = 0 .. floor(N/D)-1 (frame semantics == the trim at create_pdws_channelized.m:52-54),
//   channels k = 0..M-1 in natural FFT order (0 = DC).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef std::complex<double> cd;
static const double kPi = 3.14159265358979323846264338327950288;

extern "C" {

// ---------------------------------------------------------------------------------------------
// R1  header parser  (convert_my_iq_to_mat.m:40-102)
// ---------------------------------------------------------------------------------------------
struct orc_iq_info {
  uint32_t magic, format, header_bytes, link_speed;
  uint64_t fc_hz;
  uint32_t bw_hz, fs_sps;
  double gain_db;
  uint32_t num_samples, bit_width, spare0, bytes_per_sample;
  char board_name[17], serial_number[17], fpga_version[17], fw_version[17];
  double sample_start_time;
  uint64_t payload_offset, payload_bytes;
};

static void strip_nul(const uint8_t* src, char* dst) {  // strip(string(...),char(0)), :86-89
  int b = 0, e = 16;
  while (b < e && src[b] == 0) b++;
  while (e > b && src[e - 1] == 0) e--;
  memcpy(dst, src + b, e - b);
  dst[e - b] = 0;
}

// Returns 0 ok, -3 unknown magic (:55-56), -4 bad bit width (:96-97), -5 size mismatch (:102),
// -2 short header.
int orc_parse_header(const uint8_t* p, uint64_t file_bytes, orc_iq_info* o) {
  memset(o, 0, sizeof *o);
  if (file_bytes < 4) return -2;
  uint64_t off = 0;
  auto u32 = [&]() { uint32_t v; memcpy(&v, p + off, 4); off += 4; return v; };
  auto u64 = [&]() { uint64_t v; memcpy(&v, p + off, 8); off += 8; return v; };
  o->magic = u32();                                         // :40
  switch (o->magic) {                                       // :42-57
    case 0x00000000u: o->format = 2; break;                 // "big endian": assume latest known (:43-45)
    case 0x01010101u: o->format = 1; break;
    case 0x02020202u: o->format = 2; break;
    case 0x03030303u: o->format = 3; break;
    default: return -3;
  }
  o->header_bytes = o->format == 1 ? 104 : 112;
  if (file_bytes < o->header_bytes) return -2;
  o->link_speed = u32();                                    // :61
  o->fc_hz = o->format == 1 ? (uint64_t)u32() : u64();      // :63-68
  o->bw_hz = u32();                                         // :70
  o->fs_sps = u32();                                        // :71
  if (o->format >= 3) { float g; memcpy(&g, p + off, 4); off += 4; o->gain_db = g; }  // :73-74
  else o->gain_db = (double)u32();                          // :75-77
  o->num_samples = u32();                                   // :79
  o->bit_width = u32();                                     // :80
  if (o->format >= 2) o->spare0 = u32();                    // :82-84
  strip_nul(p + off, o->board_name); off += 16;             // :86
  strip_nul(p + off, o->serial_number); off += 16;          // :87
  strip_nul(p + off, o->fpga_version); off += 16;           // :88
  strip_nul(p + off, o->fw_version); off += 16;             // :89
  memcpy(&o->sample_start_time, p + off, 8); off += 8;      // :90
  if (o->bit_width > 0 && o->bit_width <= 8) o->bytes_per_sample = 2;          // :92-93
  else if (o->bit_width > 8 && o->bit_width <= 16) o->bytes_per_sample = 4;    // :94-95
  else return -4;                                           // :96-97
  o->payload_offset = off;
  o->payload_bytes = file_bytes - off;
  // fread(fid,[2,inf]) keeps whole pairs only; assert(length(iq) == numSamples)  (:93-95,:102)
  uint64_t pairs = o->payload_bytes / o->bytes_per_sample;
  if (pairs != o->num_samples) return -5;
  o->payload_bytes = pairs * o->bytes_per_sample;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// R2  normalise  (create_pdws_channelized.m:35-38): double(iq)/2^(bitWidth-1), row1 + 1j*row2
// out: n complex doubles, interleaved re,im
// ---------------------------------------------------------------------------------------------
void orc_unpack(const void* iq, uint64_t n, uint32_t bit_width, double* out) {
  const double max_val = std::ldexp(1.0, (int)bit_width - 1);   // :35
  if (bit_width <= 8) {
    const int8_t* s = (const int8_t*)iq;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(2 * n); i++) out[i] = (double)s[i] / max_val;
  } else {
    const int16_t* s = (const int16_t*)iq;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(2 * n); i++) out[i] = (double)s[i] / max_val;
  }
}

// ---------------------------------------------------------------------------------------------
// R4  default prototype: Kaiser-windowed sinc, cutoff fs/(2M), unity DC gain.
// dsp.Channelizer defaults NumTapsPerBand=12, StopbandAttenuation=80 (create_pdws_channelized.m:33
// uses the one-argument constructor).  Closed-source designer => our own, cross-checked against
// scipy.signal.firwin(L, 1/M, window=('kaiser', beta)) in tests.
// ---------------------------------------------------------------------------------------------
static double bessel_i0(double x) {
  double s = 1.0, t = 1.0, q = x * x / 4.0;
  for (int k = 1; k < 500; k++) { t *= q / ((double)k * k); s += t; if (t < 1e-18 * s) break; }
  return s;
}
double orc_kaiser_beta(double a) {
  if (a > 50.0) return 0.1102 * (a - 8.7);
  if (a >= 21.0) return 0.5842 * std::pow(a - 21.0, 0.4) + 0.07886 * (a - 21.0);
  return 0.0;
}
void orc_design_prototype(uint32_t M, uint32_t taps_per_band, double atten_db, double* h) {
  const int L = (int)(M * taps_per_band);
  const double beta = orc_kaiser_beta(atten_db), c = (L - 1) / 2.0, i0b = bessel_i0(beta);
  double sum = 0;
  for (int n = 0; n < L; n++) {
    double t = n - c, x = t / (double)M;
    double sinc = (t == 0.0) ? 1.0 : std::sin(kPi * x) / (kPi * x);
    double r = 2.0 * n / (L - 1) - 1.0;
    double w = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
    h[n] = sinc * w;
    sum += h[n];
  }
  for (int n = 0; n < L; n++) h[n] /= sum;
}

// ---------------------------------------------------------------------------------------------
// FFT helpers: y_k = sum_r v_r e^{+j 2 pi k r / M}.  Radix-2 for powers of two, O(M^2) otherwise
// (the reference's natural M = fs*1e-6 = 56 is not a power of two, create_pdws_channelized.m:31).
// ---------------------------------------------------------------------------------------------
struct FftPlan {
  int M; bool pow2; std::vector<cd> w; std::vector<int> rev;
  explicit FftPlan(int m) : M(m) {
    pow2 = (m & (m - 1)) == 0;
    w.resize(m);
    for (int i = 0; i < m; i++) w[i] = cd(std::cos(2 * kPi * i / m), std::sin(2 * kPi * i / m));
    if (pow2) {
      rev.resize(m);
      int lg = 0; while ((1 << lg) < m) lg++;
      for (int i = 0; i < m; i++) { int r = 0; for (int b = 0; b < lg; b++) if (i >> b & 1) r |= 1 << (lg - 1 - b); rev[i] = r; }
    }
  }
  void run(cd* a, cd* tmp) const {   // in-place on a (tmp: M scratch)
    if (!pow2) {
      for (int k = 0; k < M; k++) { cd s = 0; for (int r = 0; r < M; r++) s += a[r] * w[(int)(((int64_t)k * r) % M)]; tmp[k] = s; }
      std::copy(tmp, tmp + M, a);
      return;
    }
    for (int i = 0; i < M; i++) if (i < rev[i]) std::swap(a[i], a[rev[i]]);
    for (int len = 2; len <= M; len <<= 1) {
      int half = len >> 1, step = M / len;
      for (int i = 0; i < M; i += len)
        for (int j = 0; j < half; j++) {
          cd t = a[i + j + half] * w[j * step];
          a[i + j + half] = a[i + j] - t;
          a[i + j] += t;
        }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// R5  polyphase analysis bank.  x: n complex doubles (interleaved).  out: nrows*M complex doubles,
// row-major [row][k], natural order.  Returns rows = floor(n/D).  row0/nrows_sel select a row range
// (nrows_sel == 0 -> all) so bench/tests can compute a bounded slice of a long recording.
//   u_p[m] = sum_q h[qM+p] x[mD - qM - p];  u' = roll(u, -(mD mod M));  y[m] = M*IFFT(u')
// ---------------------------------------------------------------------------------------------
uint64_t orc_channelize_rows(const double* x, uint64_t n, uint32_t M, const double* h, uint32_t ntaps,
                             uint32_t oversample, uint64_t row0, uint64_t nrows_sel, double* out) {
  const int64_t D = M / oversample, P = ntaps / M, Mi = M;
  const uint64_t rows_total = n / (uint64_t)D;
  if (row0 > rows_total) row0 = rows_total;
  uint64_t nr = nrows_sel ? std::min<uint64_t>(nrows_sel, rows_total - row0) : rows_total - row0;
  const cd* xc = (const cd*)x;
  FftPlan plan((int)M);
#pragma omp parallel
  {
    std::vector<cd> u(M), v(M), tmp(M);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < (int64_t)nr; r++) {
      const int64_t m = (int64_t)row0 + r, t = m * D;
      for (int64_t p = 0; p < Mi; p++) {
        cd acc = 0;
        for (int64_t q = 0; q < P; q++) {
          int64_t idx = t - q * Mi - p;
          if (idx < 0) break;
          acc += h[q * Mi + p] * xc[idx];
        }
        u[p] = acc;
      }
      const int64_t sh = t % Mi;
      if (sh == 0) std::copy(u.begin(), u.end(), v.begin());
      else for (int64_t rr = 0; rr < Mi; rr++) v[rr] = u[(rr + sh) % Mi];
      plan.run(v.data(), tmp.data());
      memcpy(out + 2 * (uint64_t)r * M, v.data(), sizeof(cd) * M);
    }
  }
  return nr;
}

uint64_t orc_channelize(const double* x, uint64_t n, uint32_t M, const double* h, uint32_t ntaps,
                        uint32_t oversample, double* out) {
  return orc_channelize_rows(x, n, M, h, ntaps, oversample, 0, 0, out);
}

// Raw-payload entry (unpack fused, as the timed CPU baseline runs it): iq = int8/int16 pairs.
// Only the samples a row range needs are unpacked.  out as above.
uint64_t orc_channelize_raw(const void* iq, uint64_t n, uint32_t bit_width, uint32_t M, const double* h,
                            uint32_t ntaps, uint32_t oversample, uint64_t row0, uint64_t nrows_sel,
                            double* out) {
  const int64_t D = M / oversample, P = ntaps / M, Mi = M;
  const uint64_t rows_total = n / (uint64_t)D;
  if (row0 > rows_total) row0 = rows_total;
  uint64_t nr = nrows_sel ? std::min<uint64_t>(nrows_sel, rows_total - row0) : rows_total - row0;
  const double scale = 1.0 / std::ldexp(1.0, (int)bit_width - 1);
  const int8_t* s8 = (const int8_t*)iq; const int16_t* s16 = (const int16_t*)iq;
  const bool is8 = bit_width <= 8;
  FftPlan plan((int)M);
#pragma omp parallel
  {
    std::vector<cd> u(M), v(M), tmp(M);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < (int64_t)nr; r++) {
      const int64_t m = (int64_t)row0 + r, t = m * D;
      for (int64_t p = 0; p < Mi; p++) {
        double ar = 0, ai = 0;
        for (int64_t q = 0; q < P; q++) {
          int64_t idx = t - q * Mi - p;
          if (idx < 0) break;
          double xr, xi;
          if (is8) { xr = s8[2 * idx] * scale; xi = s8[2 * idx + 1] * scale; }
          else { xr = s16[2 * idx] * scale; xi = s16[2 * idx + 1] * scale; }
          const double hh = h[q * Mi + p];
          ar += hh * xr; ai += hh * xi;
        }
        u[p] = cd(ar, ai);
      }
      const int64_t sh = t % Mi;
      if (sh == 0) std::copy(u.begin(), u.end(), v.begin());
      else for (int64_t rr = 0; rr < Mi; rr++) v[rr] = u[(rr + sh) % Mi];
      plan.run(v.data(), tmp.data());
      memcpy(out + 2 * (uint64_t)r * M, v.data(), sizeof(cd) * M);
    }
  }
  return nr;
}

// Direct-form evaluation of the defining sum, O(rows * L * M); validates the polyphase form.
uint64_t orc_channelize_direct(const double* x, uint64_t n, uint32_t M, const double* h, uint32_t ntaps,
                               uint32_t oversample, double* out) {
  const int64_t D = M / oversample, L = ntaps;
  const uint64_t rows = n / (uint64_t)D;
  const cd* xc = (const cd*)x;
  cd* y = (cd*)out;
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < (int64_t)rows; m++)
    for (int64_t k = 0; k < (int64_t)M; k++) {
      cd acc = 0;
      for (int64_t nn = 0; nn < L; nn++) {
        int64_t idx = m * D - nn;
        if (idx < 0) break;
        double ph = 2 * kPi * (double)((k * nn) % (int64_t)M) / (double)M;
        acc += h[nn] * cd(std::cos(ph), std::sin(ph)) * xc[idx];
      }
      double ph0 = -2 * kPi * (double)((k * ((m * D) % (int64_t)M)) % (int64_t)M) / (double)M;
      y[m * (int64_t)M + k] = acc * cd(std::cos(ph0), std::sin(ph0));
    }
  return rows;
}

// R7  centerFrequencies(channelizer,fs) as used against the fftshift-ed columns
// (create_pdws_channelized.m:42,60,80): shifted column c -> (c - floor(M/2)) * fs / M.
void orc_center_frequencies(uint32_t M, double fs, double* f) {
  for (uint32_t c = 0; c < M; c++) f[c] = ((double)c - (double)(M / 2)) * fs / (double)M;
}

// ---------------------------------------------------------------------------------------------
// R6, R8-R11  PDW extraction  (create_pdws_channelized.m:60-136)
// y: nrows*M complex doubles, row-major, NATURAL channel order (the fftshift of :60 is applied here).
// ---------------------------------------------------------------------------------------------
struct orc_pdw {
  double toa_s, pw_s, freq_hz, amp, snr_db, noise_floor;
  uint32_t channel, channel_natural;
  uint64_t toa_row, end_row;
  uint32_t saturated, reserved;
};
struct orc_pdw_params {
  double snr_threshold_db, sat_level, fc_hz, fs_sps, t0;
  uint32_t reproduce_phase_bug, use_trailing_threshold;
  double trailing_snr_threshold_db;   // matlab/create_pdws.m:47 (wideband script: 3 dB)
};

static double median_of(std::vector<double>& v) {   // MATLAB median: mean of the middle two if even
  const size_t n = v.size();
  if (n == 0) return NAN;
  std::sort(v.begin(), v.end());
  return (n & 1) ? v[n / 2] : 0.5 * (v[n / 2 - 1] + v[n / 2]);
}

// D = decimation (M for the reference's critically sampled object).  Returns the PDW count; writes
// at most cap records.  noise_floor (optional): M doubles in NATURAL channel order.
uint64_t orc_pdws(const double* y, uint64_t nrows, uint32_t M, uint32_t D, const orc_pdw_params* prm,
                  orc_pdw* out, uint64_t cap, double* noise_floor) {
  const cd* yc = (const cd*)y;
  const double fs_dec = prm->fs_sps / (double)D;                          // :62
  const double thr_scale = std::pow(10.0, prm->snr_threshold_db / 10.0);  // :75 (power dB on amplitude, as written)
  const uint32_t half_up = (M + 1) / 2;                                   // fftshift: shifted c <- natural (c+ceil(M/2)) mod M
  std::vector<std::vector<orc_pdw>> per_chan(M);
  std::vector<double> bin_freqs(M);
  orc_center_frequencies(M, prm->fs_sps, bin_freqs.data());               // :42
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t bin = 0; bin < (int64_t)M; bin++) {                        // :79  (bin is 0-based here)
    const uint32_t k = (uint32_t)((bin + half_up) % M);                   // :60
    const uint32_t kph = prm->reproduce_phase_bug ? (uint32_t)(half_up % M) : k;   // :114 indexes column 1
    std::vector<double> mag(nrows);
    for (uint64_t j = 0; j < nrows; j++) mag[j] = std::abs(yc[j * M + k]);          // :67
    std::vector<double> tmp(mag);
    const double nf = median_of(tmp);                                     // :73
    if (noise_floor) noise_floor[k] = nf;
    const double thr = nf * thr_scale;                                    // :75
    // wideband script (matlab/create_pdws.m:45-47,63): a separate, lower trailing-edge threshold
    const double thr_trail = prm->use_trailing_threshold ? nf * std::pow(10.0, prm->trailing_snr_threshold_db / 10.0) : thr;
    const double fc_chan = prm->fc_hz + bin_freqs[bin];                   // :80
    bool active = false, saturated = false;                               // :82-83
    uint64_t toa = 0;
    for (uint64_t jj = 1; jj <= nrows; jj++) {                            // :85 (1-based like the script)
      const double mg = mag[jj - 1];
      if (!active) {                                                      // :87
        if (mg >= thr) { active = true; toa = jj; saturated = false; }    // :88-91
      } else if (mg <= thr_trail) {                                       // :94 (create_pdws.m:63)
        active = false;                                                   // :95
        orc_pdw r; memset(&r, 0, sizeof r);
        r.toa_s = ((double)toa / fs_dec) + prm->t0;                       // :98
        std::vector<double> seg(mag.begin() + (toa - 1), mag.begin() + jj);
        r.amp = median_of(seg);                                           // :101
        r.snr_db = 10.0 * std::log10(r.amp / nf);                         // :105
        r.pw_s = (double)(jj - toa) / fs_dec;                             // :110
        std::vector<double> pd(jj - toa);
        for (uint64_t i = toa; i < jj; i++) {                             // :114  diff(phase(toa:jj))
          const cd a = yc[(i - 1) * M + kph], b = yc[i * M + kph];
          double d = std::atan2(b.imag(), b.real()) * (180.0 / kPi) - std::atan2(a.imag(), a.real()) * (180.0 / kPi);  // :68
          if (d < -180.0) d += 360.0;                                     // :115
          if (d > 180.0) d -= 360.0;                                      // :116
          pd[i - toa] = d;
        }
        const double med = median_of(pd);                                 // :117
        r.freq_hz = fc_chan + (fs_dec / (360.0 / med));                   // :122
        r.noise_floor = nf;
        r.channel = (uint32_t)bin; r.channel_natural = k;
        r.toa_row = toa; r.end_row = jj; r.saturated = saturated ? 1u : 0u;
        per_chan[bin].push_back(r);                                       // :124-128
      } else {                                                            // :129
        const cd v = yc[(jj - 1) * M + k];
        if (std::fabs(v.real()) >= prm->sat_level || std::fabs(v.imag()) >= prm->sat_level) saturated = true;  // :130-132
      }
    }
    // a pulse still open at the end of the file is dropped (the loop just ends, :135)
  }
  uint64_t n = 0;
  for (uint32_t bin = 0; bin < M; bin++)
    for (const orc_pdw& r : per_chan[bin]) { if (n < cap && out) out[n] = r; n++; }
  return n;
}

// FSM + medians on a bare magnitude trace (hand-built known-answer tests).  Writes (toa,end) pairs,
// 1-based; returns the count.
uint64_t orc_fsm_trace(const double* mag, uint64_t n, double thr, uint64_t* toa_end, uint64_t cap) {
  bool active = false; uint64_t toa = 0, cnt = 0;
  for (uint64_t jj = 1; jj <= n; jj++) {
    if (!active) { if (mag[jj - 1] >= thr) { active = true; toa = jj; } }
    else if (mag[jj - 1] <= thr) { active = false; if (cnt < cap) { toa_end[2 * cnt] = toa; toa_end[2 * cnt + 1] = jj; } cnt++; }
  }
  return cnt;
}

double orc_median(const double* v, uint64_t n) { std::vector<double> t(v, v + n); return median_of(t); }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"
