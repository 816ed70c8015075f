// libchannelizer: C ABI implementation (handles, streaming state, launches, host pipeline).
// See include/channelizer.h for the contract and the reference lines each entry point replaces.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "chz_internal.h"
#include "chz_kernels.cuh"
#include "chz_launch.h"

namespace chzi {

static thread_local std::string g_cuda_err;

void set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s: %s (%s) at %s:%d", what, cudaGetErrorString(e), cudaGetErrorName(e), file, line);
  g_cuda_err = buf;
}

// ------------------------------------------------------------------------------------------------
// launch tables
// ------------------------------------------------------------------------------------------------
LaunchPlan plan_spans(const ::chz* h, long long nrows, int P, int groups_per_block, int blocks_per_sm,
                      int max_blocks_override) {
  LaunchPlan lp;
  const long long rows_per_phase = (nrows + h->os - 1) / h->os + 1;   // +1: a phase may start one row early (make_span)
  const long long max_blocks = max_blocks_override ? max_blocks_override : (long long)h->sm_count * blocks_per_sm;
  const long long total_groups = max_blocks * groups_per_block;
  long long sr = (rows_per_phase + total_groups * 4 - 1) / (total_groups * 4);
  sr = (sr + P - 1) / P * P;
  static const long long span_cap = std::getenv("CHZ_SPAN_CAP") ? std::atoll(std::getenv("CHZ_SPAN_CAP")) : 4096;   // tuning aid
  const long long lo = 4LL * P, hi = (span_cap / P) * P;
  if (sr < lo) {
    // Small call (a 100 ms file): at the shortest span there are between one and four spans per group, handed out
    // round-robin, so a partial last round costs a whole one (342 spans on 296 groups: two rounds for 1.16 rounds of
    // work).  Take the span length that minimises rounds x (rows + warm-up rows) instead.
    long long best = -1, best_sr = lo;
    for (long long c = lo; c <= hi && c <= 8 * lo; c += P) {
      const long long nsp = (rows_per_phase + c - 1) / c * h->os, rounds = (nsp + total_groups - 1) / total_groups;
      const long long cost = rounds * (c + P);
      if (best < 0 || cost < best) { best = cost; best_sr = c; }
      if (rounds == 1) break;                       // longer spans only add rows from here on
    }
    sr = best_sr;
  }
  if (sr > hi) sr = hi;
  lp.span_rows = (int)sr;
  lp.spans_per_phase = (rows_per_phase + sr - 1) / sr;
  const long long nspans = lp.spans_per_phase * h->os;
  long long blocks = (nspans + groups_per_block - 1) / groups_per_block;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  lp.grid = dim3((unsigned)blocks);
  return lp;
}

template <int P, bool IN16, int MT>
static int launch_fir_m(::chz* h, ChanParams prm, float2* u, cudaStream_t st) {
  const int bpb = MT ? (MT < 128 ? MT : 128) : h->fir_bpb, nbb = prm.M / bpb, groups = 128 / bpb;   // as in k_fir
  prm.bpb = bpb;
  // 128 threads x <= 128 registers: 4 blocks resident per SM; span blocks per branch block = SMs*4 / nbb
  LaunchPlan lp = plan_spans(h, prm.nrows, P, groups, 4, (h->sm_count * 4 / nbb) > 0 ? (h->sm_count * 4 / nbb) : 1);
  prm.span_rows = lp.span_rows;
  prm.spans_per_phase = lp.spans_per_phase;
  const unsigned grid = lp.grid.x * (unsigned)nbb;
  k_fir<P, IN16, MT><<<grid, 128, 0, st>>>(prm, u);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}
template <int P, bool IN16>
static int launch_fir(::chz* h, const ChanParams& prm, float2* u, cudaStream_t st) {
  if (P == 16 || P == 12) {   // the large-M sizes the split path exists for get compile-time strides
    switch (prm.M) {
      case 1024: return launch_fir_m<P, IN16, (P == 16 || P == 12) ? 1024 : 0>(h, prm, u, st);
      case 2048: return launch_fir_m<P, IN16, (P == 16 || P == 12) ? 2048 : 0>(h, prm, u, st);
      case 4096: return launch_fir_m<P, IN16, (P == 16 || P == 12) ? 4096 : 0>(h, prm, u, st);
      default: break;
    }
  }
  return launch_fir_m<P, IN16, 0>(h, prm, u, st);
}

template <int M, int ROWS, int NT>
static int launch_fft_rows_t(::chz* h, const float2* u, float2* y, long long nrows, cudaStream_t st) {
  auto kern = k_fft_rows<M, ROWS, NT>;
  const size_t smem = (size_t)(2 * ROWS * RowStride<M>::value + M) * sizeof(float2);
  static thread_local int blocks_per_sm_dev[kMaxDev] = {0};   // launch geometry is cached per device
  int& blocks_per_sm = blocks_per_sm_dev[h->device % kMaxDev];
  if (!blocks_per_sm) {
    CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CHZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NT, smem));
    blocks_per_sm = nb > 0 ? nb : 1;
  }
  long long blocks = (nrows + ROWS - 1) / ROWS;
  const long long maxb = (long long)h->sm_count * blocks_per_sm;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) return CHZ_OK;
  kern<<<(unsigned)blocks, NT, smem, st>>>(u, y, h->d_tw, nrows);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}

template <int M, int ROWS>
static int launch_fft_rows_big(::chz* h, const float2* u, float2* y, long long nrows, cudaStream_t st) {
  auto kern = k_fft_rows_big<M, ROWS>;
  const size_t smem = (size_t)(2 * ROWS * RowStride<M>::value) * sizeof(float2);
  static thread_local int blocks_per_sm_dev[kMaxDev] = {0};   // launch geometry is cached per device
  int& blocks_per_sm = blocks_per_sm_dev[h->device % kMaxDev];
  if (!blocks_per_sm) {
    CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CHZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, smem));
    blocks_per_sm = nb > 0 ? nb : 1;
  }
  long long blocks = (nrows + ROWS - 1) / ROWS;
  const long long maxb = (long long)h->sm_count * blocks_per_sm;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) return CHZ_OK;
  kern<<<(unsigned)blocks, 256, smem, st>>>(u, y, h->d_tw, nrows);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}

static int launch_fft_rows(::chz* h, const float2* u, float2* y, long long nrows, cudaStream_t st) {
  if (h->generic && h->mixed_np > 0) {
    if (nrows < 1) return CHZ_OK;
    const int M = (int)h->M;
    int rpb = 2048 / M;
    if (rpb < 1) rpb = 1;
    const size_t smem = (size_t)2 * rpb * M * sizeof(float2);
    static thread_local bool attr_set_dev[kMaxDev] = {false};
    bool& attr_set = attr_set_dev[h->device % kMaxDev];
    if (!attr_set) {
      CHZ_CUDA(cudaFuncSetAttribute(k_fft_rows_mixed, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * (int)sizeof(float2)));
      attr_set = true;
    }
    MixedPlan plan;
    plan.np = h->mixed_np;
    for (int i = 0; i < 12; i++) plan.r[i] = h->mixed_r[i];
    long long blocks = (nrows + rpb - 1) / rpb;
    const long long maxb = (long long)h->sm_count * 6;
    if (blocks > maxb) blocks = maxb;
    k_fft_rows_mixed<<<(unsigned)blocks, 256, smem, st>>>(u, y, h->d_tw, M, nrows, rpb, plan);
    h->launches++;
    CHZ_CUDA(cudaGetLastError());
    return CHZ_OK;
  }
  if (h->generic) {
    if (nrows < 1) return CHZ_OK;
    const size_t smem = (size_t)2 * h->M * sizeof(float2);
    static thread_local bool attr_set_dev[kMaxDev] = {false};
    bool& attr_set = attr_set_dev[h->device % kMaxDev];
    if (!attr_set) {
      CHZ_CUDA(cudaFuncSetAttribute(k_dft_rows_any, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * (int)sizeof(float2)));
      attr_set = true;
    }
    long long blocks = nrows < (long long)h->sm_count * 8 ? nrows : (long long)h->sm_count * 8;
    k_dft_rows_any<<<(unsigned)blocks, 256, smem, st>>>(u, y, h->d_tw, (int)h->M, nrows);
    h->launches++;
    CHZ_CUDA(cudaGetLastError());
    return CHZ_OK;
  }
  switch (h->M) {
    case 8: return launch_fft_rows_t<8, 256, 256>(h, u, y, nrows, st);
    case 16: return launch_fft_rows_t<16, 128, 256>(h, u, y, nrows, st);
    case 32: return launch_fft_rows_t<32, 64, 256>(h, u, y, nrows, st);
    case 64: return launch_fft_rows_t<64, 32, 256>(h, u, y, nrows, st);
    case 56: return launch_fft_rows_t<56, 32, 256>(h, u, y, nrows, st);
    case 560: return launch_fft_rows_t<560, 4, 256>(h, u, y, nrows, st);
    case 128: return launch_fft_rows_t<128, 16, 256>(h, u, y, nrows, st);
    case 256: return launch_fft_rows_t<256, 16, 256>(h, u, y, nrows, st);
    case 512: return launch_fft_rows_big<512, 8>(h, u, y, nrows, st);
    case 1024: return launch_fft_rows_big<1024, 4>(h, u, y, nrows, st);
    case 2048: return launch_fft_rows_big<2048, 2>(h, u, y, nrows, st);
    case 4096: return launch_fft_rows_big<4096, 1>(h, u, y, nrows, st);
    default: return CHZ_EINVAL;
  }
}

template <bool IN16>
static int launch_fir_dispatch(::chz* h, const ChanParams& prm, float2* u, cudaStream_t st) {
  switch ((h->generic && h->fir_bpb == 0) ? 0u : h->P) {
    case 4: return launch_fir<4, IN16>(h, prm, u, st);
    case 8: return launch_fir<8, IN16>(h, prm, u, st);
    case 12: return launch_fir<12, IN16>(h, prm, u, st);
    case 16: return launch_fir<16, IN16>(h, prm, u, st);
    case 24: return launch_fir<24, IN16>(h, prm, u, st);
    case 32: return launch_fir<32, IN16>(h, prm, u, st);
    default: {
      const long long total = prm.nrows * prm.M;
      long long blocks = (total + 255) / 256;
      const long long maxb = (long long)h->sm_count * 8;
      if (blocks > maxb) blocks = maxb;
      k_fir_any<IN16><<<(unsigned)blocks, 256, 0, st>>>(prm, (int)h->P, u);
      h->launches++;
      CHZ_CUDA(cudaGetLastError());
      return CHZ_OK;
    }
  }
}

static bool fused_available(const ::chz* h) {
  return !h->generic && h->M >= 8 && h->M <= 560 && (h->P == 8 || h->P == 12 || h->P == 16);
}

static int ensure_taps(::chz* h, uint32_t bw) {
  if (h->d_taps[bw]) return CHZ_OK;
  std::vector<float> scaled(h->L);
  const float s = std::ldexp(1.0f, -(int)(bw - 1));   // exact power of two (create_pdws_channelized.m:35-37)
  for (uint32_t i = 0; i < h->L; i++) scaled[i] = h->taps[i] * s;
  CHZ_CUDA(cudaMalloc(&h->d_taps[bw], sizeof(float) * h->L));
  CHZ_CUDA(cudaMemcpy(h->d_taps[bw], scaled.data(), sizeof(float) * h->L, cudaMemcpyHostToDevice));
  return CHZ_OK;
}

// One launch (or FIR + FFT pair) over `nsamp` new device-resident samples; updates the stream state.
static int run_chunk(::chz* h, const void* iq_dev, uint64_t nsamp, uint32_t bw, float2* out_dev,
                     uint64_t* nrows_out, cudaStream_t st) {
  NvtxRange nvtx_range("chz:channelize");
  const bool in16 = bw > 8;
  const size_t bps = in16 ? 4 : 2;
  const uint64_t rows_new = (h->consumed + nsamp) / h->D - h->rows_done;
  int rc = ensure_taps(h, bw);
  if (rc) return rc;
  if (rows_new > 0) {
    ChanParams prm;
    memset(&prm, 0, sizeof prm);
    prm.in = iq_dev; prm.hist = h->d_hist[h->hist_cur];
    prm.in_base = (long long)h->consumed; prm.n_in = (long long)nsamp; prm.hist_base = (long long)h->hist_base;
    prm.taps = h->d_taps[bw]; prm.tw = h->d_tw; prm.out = out_dev;
    prm.row_base = (long long)h->rows_done; prm.nrows = (long long)rows_new;
    prm.M = (int)h->M; prm.D = (int)h->D; prm.os = (int)h->os;
    const bool fused = fused_available(h) && h->force_path != 2;
    // The cluster kernel is opt-in (CHZ_OPT_FORCE_PATH = 3): measured on B200 it is slower than the
    // split path (cfg4: 27.8 % vs 39.3 % of the HBM roofline; only 120 of 148 SMs host 8-CTA clusters and
    // each tile serialises FIR -> release fence -> cluster barrier -> L2 reads -> 3 FFT passes).
    const bool cl_path = h->force_path == 3 || h->force_path == 7 || h->force_path == 8 || h->force_path == 9;
    const int cl_tpc = (h->force_path == 7 || h->force_path == 8) ? 256 : 512;
    const bool cluster = cl_path && cluster_available(h, cl_tpc);
    if (cl_path && !cluster) return CHZ_EINVAL;
    const bool ws = ws_available(h) && h->force_path == 4;
    if (h->force_path == 4 && !ws) return CHZ_EINVAL;
    if (h->force_path == 1 && !fused) return CHZ_EINVAL;
    const bool dit2 = dit2_available(h) && h->force_path == 5;
    if (h->force_path == 5 && !dit2) return CHZ_EINVAL;
    const bool dsm = dsm_available(h) && h->force_path == 10;
    if (h->force_path == 10 && !dsm) return CHZ_EINVAL;
    const bool pipe = pipe_available(h) && h->force_path == 6;
    const bool ringp = ring_available(h) && (h->force_path == 0 || h->force_path == 11);
    if (h->force_path == 11 && !ringp) return CHZ_EINVAL;
    if (h->force_path == 6 && !pipe) return CHZ_EINVAL;
    if (ringp) {
      rc = launch_ring_any(h, prm, in16, st);
      if (rc) return rc == 1 ? CHZ_EINVAL : rc;
    } else if (dsm) {
      rc = launch_dsm_any(h, prm, in16, st);
      if (rc) return rc == 1 ? CHZ_EINVAL : rc;
    } else if (pipe) {
      rc = launch_pipe_any(h, prm, in16, st);
      if (rc) return rc == 1 ? CHZ_EINVAL : rc;
    } else if (dit2) {
      rc = launch_dit2_any(h, prm, in16, st);
      if (rc) return rc == 1 ? CHZ_EINVAL : rc;
    } else if (ws) {
      rc = launch_ws_any(h, prm, in16, st);
      if (rc) return rc == 1 ? CHZ_EINVAL : rc;
    } else     if (cluster) {
      rc = launch_cluster_any(h, prm, in16, h->force_path, st);
      if (rc == 1) return CHZ_EINVAL;
      if (rc) return rc;
    } else if (fused) {
      rc = launch_fused_any(h, prm, in16, st);
      if (rc == 1) return CHZ_EINVAL;
      if (rc) return rc;
    } else {
      // split path: FIR rows -> the output buffer itself, then the row FFT over it IN PLACE: 4 + 8 + 8 + 8 B per
      // sample through DRAM.  By default the whole call is one FIR launch and one FFT launch.  CHZ_SPLIT_CHUNK_MB
      // walks the rows in chunks small enough for the FIR output to stay in the 126 MB L2 until the FFT reads it;
      // measured slower on B200 (the small launches lose more to ramps and tails than the L2 hits give back,
      // DESIGN.md section 4), so it stays a tuning aid.
      long long chunk_rows = (long long)(h->split_chunk_bytes / ((uint64_t)h->M * sizeof(float2)));
      chunk_rows &= ~1LL;                       // even: row pairs stay aligned to global parity
      if (chunk_rows < 2) chunk_rows = 2;
      const long long nchunks = ((long long)rows_new + chunk_rows - 1) / chunk_rows;
      const bool overlap = nchunks > 2 && h->split_overlap;
      if (overlap && !h->ev_fir[0]) {
        for (int i = 0; i < 4; i++) {
          CHZ_CUDA(cudaEventCreateWithFlags(&h->ev_fir[i], cudaEventDisableTiming));
          CHZ_CUDA(cudaEventCreateWithFlags(&h->ev_fft[i], cudaEventDisableTiming));
        }
      }
      // overlap mode: FIR(i+1) on the main stream runs concurrently with FFT(i) on a side stream, at most two
      // chunks ahead, so the two small kernels fill each other's launch gaps and tails
      cudaStream_t sb = overlap ? h->s_h2d : st;
      long long ci = 0;
      for (long long r0 = 0; r0 < (long long)rows_new; r0 += chunk_rows, ci++) {
        const long long nr = std::min<long long>(chunk_rows, (long long)rows_new - r0);
        ChanParams cp = prm;
        cp.row_base = prm.row_base + r0;
        cp.nrows = nr;
        float2* dst = out_dev + r0 * (long long)h->M;
        cp.out = dst;
        if (overlap && ci >= 2) CHZ_CUDA(cudaStreamWaitEvent(st, h->ev_fft[(ci - 2) & 3], 0));
        rc = in16 ? launch_fir_dispatch<true>(h, cp, dst, st) : launch_fir_dispatch<false>(h, cp, dst, st);
        if (rc) return rc;
        if (overlap) {
          CHZ_CUDA(cudaEventRecord(h->ev_fir[ci & 3], st));
          CHZ_CUDA(cudaStreamWaitEvent(sb, h->ev_fir[ci & 3], 0));
        }
        rc = launch_fft_rows(h, dst, dst, nr, sb);
        if (rc) return rc;
        if (overlap) CHZ_CUDA(cudaEventRecord(h->ev_fft[ci & 3], sb));
      }
      if (overlap) {   // the main stream continues only after the last FFTs
        CHZ_CUDA(cudaStreamWaitEvent(st, h->ev_fft[(ci - 1) & 3], 0));
        if (ci >= 2) CHZ_CUDA(cudaStreamWaitEvent(st, h->ev_fft[(ci - 2) & 3], 0));
      }
    }
  }
  // history for the next call: everything from the oldest sample the next row can touch
  const uint64_t end = h->consumed + nsamp;
  const uint64_t next_newest = (h->rows_done + rows_new) * (uint64_t)h->D;
  const uint64_t need_from = next_newest >= (uint64_t)(h->L - 1) ? next_newest - (h->L - 1) : 0;
  const uint64_t from = need_from < end ? need_from : end;
  char* dst = (char*)h->d_hist[1 - h->hist_cur];
  uint64_t written = 0;
  if (from < h->consumed) {   // part still inside the old history
    const uint64_t a = from > h->hist_base ? from : h->hist_base;
    const uint64_t n_old = h->consumed - a;
    if (n_old)
      CHZ_CUDA(cudaMemcpyAsync(dst, (const char*)h->d_hist[h->hist_cur] + (a - h->hist_base) * bps, n_old * bps,
                               cudaMemcpyDeviceToDevice, st));
    written = n_old;
    h->hist_base = a;
    if (nsamp) CHZ_CUDA(cudaMemcpyAsync(dst + written * bps, iq_dev, nsamp * bps, cudaMemcpyDeviceToDevice, st));
    written += nsamp;
  } else {
    const uint64_t n_new = end - from;
    if (n_new)
      CHZ_CUDA(cudaMemcpyAsync(dst, (const char*)iq_dev + (from - h->consumed) * bps, n_new * bps,
                               cudaMemcpyDeviceToDevice, st));
    written = n_new;
    h->hist_base = from;
  }
  h->hist_len = written;
  h->hist_cur = 1 - h->hist_cur;
  h->consumed = end;
  h->rows_done += rows_new;
  if (nrows_out) *nrows_out = rows_new;
  return CHZ_OK;
}

static int ensure_store(::chz* h, uint64_t rows_needed) {
  if (rows_needed <= h->store_cap) return CHZ_OK;
  uint64_t cap = h->store_cap * 2;
  if (cap < rows_needed) cap = rows_needed;
  float2* nb = nullptr;
  {
    cudaError_t e = cudaMalloc(&nb, cap * h->M * sizeof(float2));
    if (e == cudaErrorMemoryAllocation && cap > rows_needed) {   // doubling does not fit: take exactly what is needed
      cudaGetLastError();
      cap = rows_needed;
      e = cudaMalloc(&nb, cap * h->M * sizeof(float2));
    }
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return CHZ_ENOMEM; }
    CHZ_CUDA(e);
  }
  if (h->store_rows) {
    CHZ_CUDA(cudaMemcpyAsync(nb, h->d_store, h->store_rows * h->M * sizeof(float2), cudaMemcpyDeviceToDevice, h->stream));
    CHZ_CUDA(cudaStreamSynchronize(h->stream));
  }
  if (h->d_store) CHZ_CUDA(cudaFree(h->d_store));
  h->d_store = nb;
  h->store_cap = cap;
  return CHZ_OK;
}

static int check_stream_args(::chz* h, uint64_t nsamp, uint32_t bw) {
  if (bw == 0 || bw > 16) return CHZ_EBITWIDTH;
  if (h->consumed + nsamp > (1ULL << 62)) return CHZ_EINVAL;
  if (h->bit_width && h->bit_width != bw) return CHZ_ESTATE;   // one sample format per stream; chz_reset to change
  if (h->poisoned) return CHZ_ESTATE;                           // an earlier call failed half way: chz_reset first
  return CHZ_OK;
}

}  // namespace chzi

using namespace chzi;

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* chz_strerror(int code) {
  switch (code) {
    case CHZ_OK: return "ok";
    case CHZ_EINVAL: return "invalid argument";
    case CHZ_EIO: return "I/O error";
    case CHZ_EFORMAT: return "Unsupported endianness (unknown .iq magic)";
    case CHZ_EBITWIDTH: return "Unsupported bit width";
    case CHZ_ESIZE: return "payload length does not match numSamples";
    case CHZ_ENOMEM: return "out of memory";
    case CHZ_ECUDA: return "CUDA error";
    case CHZ_ENODEVICE: return "no usable sm_100 GPU (libchannelizer has no CPU path)";
    case CHZ_ECAPACITY: return "output buffer too small";
    case CHZ_ESTATE: return "call not valid in the current state";
    default: return "unknown error";
  }
}
const char* chz_last_cuda_error(void) { return g_cuda_err.c_str(); }
int chz_abi_version(void) { return CHZ_ABI_VERSION; }

int chz_create(uint32_t M, const float* taps, uint32_t ntaps, uint32_t oversample, chz_t** out) {
  if (!out) return CHZ_EINVAL;
  *out = nullptr;
  if (M < 1 || M > 4096) return CHZ_EINVAL;
  if (oversample != 1 && oversample != 2) return CHZ_EINVAL;
  if (M % oversample != 0) return CHZ_EINVAL;
  if (taps && (ntaps == 0 || ntaps % M != 0 || ntaps / M > 32)) return CHZ_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return CHZ_ENODEVICE; }
  int dev = 0;
  CHZ_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CHZ_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return CHZ_ENODEVICE;   // kernels are built for sm_100a only
  chz* h = new (std::nothrow) chz;
  if (!h) return CHZ_ENOMEM;
  h->device = dev;
  h->sm_count = prop.multiProcessorCount;
  h->M = M; h->os = oversample; h->D = M / oversample;
  // no radix plan: direct FIR + O(M^2) row DFT kernels.  56 = 8*7 and 560 = 16*5*7, the reference's own
  // channel counts (create_pdws_channelized.m:31, generate_channelized_training_iq.m:95-96), have plans.
  h->generic = (M < 8 || (M & (M - 1)) != 0) && M != 56 && M != 560;
  if (h->generic && M >= 2) {
    // run-time mixed-radix plan when M = 2^a 3^b 5^c 7^d: 16s first, then 8/4/2, then 7, 5, 3 (larger radices first)
    uint32_t m = M;
    int c2 = 0, c3 = 0, c5 = 0, c7 = 0;
    while (m % 2 == 0) { m /= 2; c2++; }
    while (m % 3 == 0) { m /= 3; c3++; }
    while (m % 5 == 0) { m /= 5; c5++; }
    while (m % 7 == 0) { m /= 7; c7++; }
    if (m == 1) {
      int np = 0;
      while (c2 >= 4) { h->mixed_r[np++] = 16; c2 -= 4; }
      if (c2 == 3) h->mixed_r[np++] = 8;
      if (c2 == 2) h->mixed_r[np++] = 4;
      if (c2 == 1) h->mixed_r[np++] = 2;
      for (int i = 0; i < c7; i++) h->mixed_r[np++] = 7;
      for (int i = 0; i < c5; i++) h->mixed_r[np++] = 5;
      for (int i = 0; i < c3; i++) h->mixed_r[np++] = 3;
      std::sort(h->mixed_r, h->mixed_r + np, [](int a, int b) { return a > b; });
      h->mixed_np = np;
    }
  }
  // register-window FIR with run-time M: M branches per 128-thread block when they fit, else the largest
  // divisor of M that does (none >= 32: the direct FIR kernel)
  if (M <= 128) h->fir_bpb = (int)M;
  else
    for (uint32_t d = 128; d >= 32; d--)
      if (M % d == 0) { h->fir_bpb = (int)d; break; }
  if (taps) {
    h->taps.assign(taps, taps + ntaps);
  } else {
    h->taps.resize((size_t)M * 12);   // dsp.Channelizer defaults: 12 taps per band, 80 dB
    chz_design_prototype(M, 12, 80.0, h->taps.data());
  }
  h->L = (uint32_t)h->taps.size();
  h->P = h->L / M;
  auto fail = [&](int rc) { chz_destroy(h); return rc; };
#define CHZ_TRY(call) do { if ((call) != cudaSuccess) { set_cuda_error(cudaGetLastError(), #call, __FILE__, __LINE__); return fail(CHZ_ECUDA); } } while (0)
  CHZ_TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  CHZ_TRY(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
  CHZ_TRY(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  for (int i = 0; i < 2; i++) {
    CHZ_TRY(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
    CHZ_TRY(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
    CHZ_TRY(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
  }
  // Inter-pass twiddles of the Stockham plan (same radices as Plan<M> on the device), laid out per pass
  // as entry (q-1)*NS + k = W_{NS R}^{q k} = e^{+j 2 pi q k / (NS R)} so a warp reads consecutive k.
  std::vector<float2> tw(M, make_float2(1.f, 0.f));
  if (h->generic && h->mixed_np == 0) {   // plain table W_M^i = e^{+j 2 pi i / M} for the O(M^2) DFT
    for (uint32_t i = 0; i < M; i++) {
      const double a = 2.0 * 3.14159265358979323846264338327950288 * (double)i / (double)M;
      tw[i] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  } else {
    int r[12] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
    int np = 3;
    if (h->generic) {
      np = h->mixed_np;
      for (int i = 0; i < np; i++) r[i] = h->mixed_r[i];
    } else {
      switch (M) {
        case 8: r[0] = 8; break;              case 16: r[0] = 16; break;
        case 32: r[0] = 8; r[1] = 4; break;   case 64: r[0] = 8; r[1] = 8; break;
        case 128: r[0] = 16; r[1] = 8; break; case 256: r[0] = 16; r[1] = 16; break;
        case 512: r[0] = 16; r[1] = 8; r[2] = 4; break;   case 1024: r[0] = 16; r[1] = 8; r[2] = 8; break;
        case 2048: r[0] = 16; r[1] = 16; r[2] = 8; break; case 4096: r[0] = 16; r[1] = 16; r[2] = 16; break;
        case 56: r[0] = 8; r[1] = 7; break;               case 560: r[0] = 16; r[1] = 5; r[2] = 7; break;
      }
    }
    size_t off = 0;
    int ns = r[0];
    for (int pass = 1; pass < np && r[pass] > 1; pass++) {
      const int R = r[pass];
      for (int q = 1; q < R; q++)
        for (int k = 0; k < ns; k++) {
          const double a = 2.0 * 3.14159265358979323846264338327950288 * (double)q * (double)k / ((double)ns * R);
          tw[off + (size_t)(q - 1) * ns + k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
      off += (size_t)(R - 1) * ns;
      ns *= R;
    }
  }
  {
    std::vector<float2> twn(M);
    for (uint32_t i = 0; i < M; i++) {
      const double a = 2.0 * 3.14159265358979323846264338327950288 * (double)i / (double)M;
      twn[i] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    CHZ_TRY(cudaMalloc(&h->d_twn, sizeof(float2) * M));
    CHZ_TRY(cudaMemcpy(h->d_twn, twn.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
  }
  if (const char* e = std::getenv("CHZ_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(e));   // tuning aid: 32 / 64 / 128
  if (const char* e = std::getenv("CHZ_PDW_GRAPH")) h->pdw_use_graph = std::atoi(e) != 0;
  if (const char* e = std::getenv("CHZ_RING_UNPACK")) h->ring_unpack = std::atoi(e);          // tuning aids
  if (const char* e = std::getenv("CHZ_RING_VARIANT")) h->ring_variant = std::atoi(e);
  if (const char* e = std::getenv("CHZ_RING_DBG")) h->ring_dbg = std::atoi(e);
  if (const char* e = std::getenv("CHZ_RING_MIN_STEPS")) { const int v = std::atoi(e); if (v > 0) h->ring_min_steps = v; }
  CHZ_TRY(cudaMalloc(&h->d_tw, sizeof(float2) * M));
  CHZ_TRY(cudaMemcpy(h->d_tw, tw.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
  if (const char* e = std::getenv("CHZ_SPLIT_OVERLAP")) h->split_overlap = std::atoi(e) != 0;   // tuning aid
  if (const char* e = std::getenv("CHZ_PIPE_SPAN_ROWS")) { const int v = std::atoi(e); if (v > 0) h->pipe_span_rows = v; }   // tuning aids
  if (const char* e = std::getenv("CHZ_PIPE_BLOCKS")) { const int v = std::atoi(e); if (v > 0) h->pipe_blocks = v; }
  if (const char* e = std::getenv("CHZ_PIPE_LAG")) { const int v = std::atoi(e); if (v > 0) h->pipe_lag = v; }
  if (const char* e = std::getenv("CHZ_SPLIT_CHUNK_MB")) {   // tuning aid
    const long v = std::atol(e);
    if (v > 0) h->split_chunk_bytes = (uint64_t)v << 20;
  }
  h->hist_cap = (uint64_t)h->L + h->D + 16;
  for (int i = 0; i < 2; i++) {
    CHZ_TRY(cudaMalloc(&h->d_hist[i], h->hist_cap * 4));
    CHZ_TRY(cudaMemset(h->d_hist[i], 0, h->hist_cap * 4));
  }
#undef CHZ_TRY
  *out = h;
  return CHZ_OK;
}

void chz_destroy(chz_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  for (int i = 0; i < 17; i++) if (h->d_taps[i]) cudaFree(h->d_taps[i]);
  if (h->d_tw) cudaFree(h->d_tw);
  if (h->d_twn) cudaFree(h->d_twn);
  for (int i = 0; i < 2; i++) {
    if (h->d_hist[i]) cudaFree(h->d_hist[i]);
    if (h->d_in[i]) cudaFree(h->d_in[i]);
    if (h->d_out[i]) cudaFree(h->d_out[i]);
    if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
    if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
    if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
  }
  if (h->d_store) cudaFree(h->d_store);
  if (h->d_u) cudaFree(h->d_u);
  h->cluster_ring.release();
  h->pipe_ctrl.release();
  for (int i = 0; i < 4; i++) { if (h->ev_fir[i]) cudaEventDestroy(h->ev_fir[i]); if (h->ev_fft[i]) cudaEventDestroy(h->ev_fft[i]); }
  for (chzi::Scratch* sc : {&h->pdw_hist, &h->pdw_sel, &h->pdw_thr, &h->pdw_cnt, &h->pdw_ev, &h->pdw_pin, &h->pdw_pout, &h->pdw_code, &h->pdw_nf, &h->pdw_fast}) sc->release();
  if (h->pdw_stage_host) cudaFreeHost(h->pdw_stage_host);
  if (h->pdw_graph) cudaGraphExecDestroy(h->pdw_graph);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  delete h;
}

int chz_reset(chz_t* h) {
  if (!h) return CHZ_EINVAL;
  h->consumed = 0; h->rows_done = 0; h->bit_width = 0;
  h->hist_base = 0; h->hist_len = 0;
  h->store_rows = 0;
  h->poisoned = false;
  h->pdws.clear(); h->noise_floor.clear();
  return CHZ_OK;
}

int chz_set_stream(chz_t* h, void* cuda_stream) {
  if (!h) return CHZ_EINVAL;
  // NULL is CUDA's default stream, exactly as for any cudaStream_t argument; (void*)-1 goes back to the
  // handle's own non-blocking stream.
  h->stream = cuda_stream == (void*)(intptr_t)-1 ? h->own_stream : (cudaStream_t)cuda_stream;
  return CHZ_OK;
}

int chz_set_option(chz_t* h, int opt, int64_t value) {
  if (!h) return CHZ_EINVAL;
  switch (opt) {
    case CHZ_OPT_RETAIN: h->retain = value != 0; return CHZ_OK;
    case CHZ_OPT_CHUNK_ROWS: if (value < 0) return CHZ_EINVAL; h->chunk_rows = value; return CHZ_OK;
    case CHZ_OPT_PDW_EVENT_PATH: h->pdw_event_path = value != 0; return CHZ_OK;
    case CHZ_OPT_FORCE_PATH: {
      bool ok = false;   // a path this build does not contain, or that has no kernel for the handle's (M, taps), is refused here
      switch (value) {
        case 0: case 2: ok = true; break;
        case 1: ok = fused_available(h); break;
        case 3: case 9: ok = cluster_available(h, 512); break;
        case 7: case 8: ok = cluster_available(h, 256); break;
        case 4: ok = ws_available(h); break;
        case 5: ok = dit2_available(h); break;
        case 6: ok = pipe_available(h); break;
        case 10: ok = dsm_available(h); break;
        case 11: ok = ring_available(h); break;
        default: break;
      }
      if (!ok) return CHZ_EINVAL;
      h->force_path = (int)value;
      return CHZ_OK;
    }
    default: return CHZ_EINVAL;
  }
}

uint32_t chz_num_channels(const chz_t* h) { return h ? h->M : 0; }
uint32_t chz_num_taps(const chz_t* h) { return h ? h->L : 0; }
uint32_t chz_decimation(const chz_t* h) { return h ? h->D : 0; }
int chz_get_taps(const chz_t* h, float* taps, uint32_t cap) {
  if (!h || !taps) return CHZ_EINVAL;
  if (cap < h->L) return CHZ_ECAPACITY;
  memcpy(taps, h->taps.data(), sizeof(float) * h->L);
  return CHZ_OK;
}
uint64_t chz_rows_for(const chz_t* h, uint64_t nsamp) {
  return h ? (h->consumed + nsamp) / h->D - h->rows_done : 0;
}
uint64_t chz_kernel_launches(const chz_t* h) { return h ? h->launches : 0; }

double chz_channel_freq(const chz_t* h, uint32_t k, double fs) {
  if (!h || k >= h->M) return NAN;
  const double step = fs / (double)h->M;
  return k < (h->M + 1) / 2 ? k * step : ((double)k - (double)h->M) * step;
}

int chz_process_dev(chz_t* h, const void* iq_dev, uint64_t nsamp, uint32_t bit_width, chz_cf32* out_dev,
                    uint64_t out_cap_rows, uint64_t* nrows) {
  if (!h || (!iq_dev && nsamp)) return CHZ_EINVAL;
  int rc = check_stream_args(h, nsamp, bit_width);
  if (rc) return rc;
  const uint64_t need = chz_rows_for(h, nsamp);
  if (nrows) *nrows = need;
  if (need > out_cap_rows) return CHZ_ECAPACITY;
  if (need && !out_dev) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  h->bit_width = bit_width;
  return run_chunk(h, iq_dev, nsamp, bit_width, (float2*)out_dev, nrows, h->stream);
}

int chz_synchronize(chz_t* h) {
  if (!h) return CHZ_EINVAL;
  CHZ_CUDA(cudaStreamSynchronize(h->stream));
  return CHZ_OK;
}

int chz_process(chz_t* h, const void* iq, uint64_t nsamp, uint32_t bit_width, chz_cf32* out,
                uint64_t out_cap_rows, uint64_t* nrows) {
  if (!h || (!iq && nsamp)) return CHZ_EINVAL;
  int rc = check_stream_args(h, nsamp, bit_width);
  if (rc) return rc;
  const uint64_t need = chz_rows_for(h, nsamp);
  if (nrows) *nrows = need;
  if (out && need > out_cap_rows) return CHZ_ECAPACITY;
  if (!out && !h->retain && need) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  NvtxRange nvtx_range("chz:process(host buffers)");
  h->bit_width = bit_width;
  const size_t bps = bit_width > 8 ? 4 : 2;
  const uint64_t D = h->D, M = h->M;
  if (nsamp == 0) { if (nrows) *nrows = 0; return CHZ_OK; }
  // chunk geometry: whole frames per chunk so each chunk yields a fixed number of rows
  uint64_t crow = h->chunk_rows > 0 ? (uint64_t)h->chunk_rows : (8ULL << 20) / D;   // ~8 Mi samples
  if (crow < 1) crow = 1;
  const uint64_t csamp = crow * D;
  if (h->in_cap_bytes < (csamp + D) * bps) {
    for (int i = 0; i < 2; i++) {
      if (h->d_in[i]) CHZ_CUDA(cudaFree(h->d_in[i]));
      h->d_in[i] = nullptr;
      CHZ_CUDA(cudaMalloc(&h->d_in[i], (csamp + D) * bps));
    }
    h->in_cap_bytes = (csamp + D) * bps;
  }
  if (h->retain) {
    rc = ensure_store(h, h->store_rows + need);
    if (rc) return rc;
  } else if (h->out_cap_rows < crow + 2) {
    for (int i = 0; i < 2; i++) {
      if (h->d_out[i]) CHZ_CUDA(cudaFree(h->d_out[i]));
      h->d_out[i] = nullptr;
      CHZ_CUDA(cudaMalloc(&h->d_out[i], (crow + 2) * M * sizeof(float2)));
    }
    h->out_cap_rows = crow + 2;
  }
  cudaStream_t sc = h->stream;
  uint64_t done = 0, rows_out = 0;
  int c = 0;
  // A failure in the middle of the pipeline leaves copies in flight that reference the caller's buffers and a
  // stream state that has advanced for some chunks only: drain the three streams before returning and mark the
  // handle so that nothing but chz_reset is accepted afterwards.
  auto pipeline = [&]() -> int {
    // make the side streams wait for whatever is already queued on the compute stream
    CHZ_CUDA(cudaEventRecord(h->ev_comp[0], sc));
    CHZ_CUDA(cudaEventRecord(h->ev_comp[1], sc));
    CHZ_CUDA(cudaEventRecord(h->ev_d2h[0], h->s_d2h));
    CHZ_CUDA(cudaEventRecord(h->ev_d2h[1], h->s_d2h));
    while (done < nsamp) {
      const int b = c & 1;
      const uint64_t n = (nsamp - done) < csamp ? (nsamp - done) : csamp;
      // H2D of chunk c may start once the kernel that last read d_in[b] (chunk c-2) is done
      CHZ_CUDA(cudaStreamWaitEvent(h->s_h2d, h->ev_comp[b], 0));
      if (n) CHZ_CUDA(cudaMemcpyAsync(h->d_in[b], (const char*)iq + done * bps, n * bps, cudaMemcpyHostToDevice, h->s_h2d));
      CHZ_CUDA(cudaEventRecord(h->ev_h2d[b], h->s_h2d));
      CHZ_CUDA(cudaStreamWaitEvent(sc, h->ev_h2d[b], 0));
      float2* dst;
      if (h->retain) dst = h->d_store + h->store_rows * M;
      else { dst = h->d_out[b]; CHZ_CUDA(cudaStreamWaitEvent(sc, h->ev_d2h[b], 0)); }
      uint64_t r = 0;
      const int rcc = run_chunk(h, h->d_in[b], n, bit_width, dst, &r, sc);
      if (rcc) return rcc;
      CHZ_CUDA(cudaEventRecord(h->ev_comp[b], sc));
      if (h->retain) h->store_rows += r;
      if (out && r) {
        CHZ_CUDA(cudaStreamWaitEvent(h->s_d2h, h->ev_comp[b], 0));
        CHZ_CUDA(cudaMemcpyAsync(out + rows_out * M, dst, r * M * sizeof(float2), cudaMemcpyDeviceToHost, h->s_d2h));
        CHZ_CUDA(cudaEventRecord(h->ev_d2h[b], h->s_d2h));
      }
      rows_out += r;
      done += n;
      c++;
    }
    CHZ_CUDA(cudaStreamSynchronize(h->s_h2d));
    CHZ_CUDA(cudaStreamSynchronize(sc));
    CHZ_CUDA(cudaStreamSynchronize(h->s_d2h));
    return CHZ_OK;
  };
  rc = pipeline();
  if (rc) {
    cudaStreamSynchronize(h->s_h2d); cudaStreamSynchronize(sc); cudaStreamSynchronize(h->s_d2h);
    h->poisoned = true;
    if (nrows) *nrows = 0;
    return rc;
  }
  if (nrows) *nrows = rows_out;
  return CHZ_OK;
}

int chz_unpack_dev(const void* iq_dev, uint64_t nsamp, uint32_t bit_width, chz_cf32* out_dev, void* cuda_stream) {
  if ((!iq_dev || !out_dev) && nsamp) return CHZ_EINVAL;
  if (bit_width == 0 || bit_width > 16) return CHZ_EBITWIDTH;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return CHZ_ENODEVICE; }
  if (!nsamp) return CHZ_OK;
  const float scale = std::ldexp(1.0f, -(int)(bit_width - 1));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const bool in16 = bit_width > 8;
  const bool aligned = (((uintptr_t)iq_dev) & (in16 ? 7u : 3u)) == 0 && (((uintptr_t)out_dev) & 15u) == 0;
  const long long n2 = aligned ? (long long)(nsamp / 2) : 0;     // sample pairs through the two-per-thread kernel
  if (n2) {
    long long blocks = (n2 + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (in16) k_unpack<true><<<(unsigned)blocks, 256, 0, st>>>(iq_dev, n2, scale, (float4*)out_dev);
    else k_unpack<false><<<(unsigned)blocks, 256, 0, st>>>(iq_dev, n2, scale, (float4*)out_dev);
    CHZ_CUDA(cudaGetLastError());
  }
  const long long rest = (long long)nsamp - 2 * n2;              // everything if misaligned, else 0 or 1 sample
  if (rest) {
    long long blocks = (rest + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const char* src = (const char*)iq_dev + (size_t)(2 * n2) * (in16 ? 4 : 2);
    float2* dst = (float2*)out_dev + 2 * n2;
    if (in16) k_unpack1<true><<<(unsigned)blocks, 256, 0, st>>>(src, rest, scale, dst);
    else k_unpack1<false><<<(unsigned)blocks, 256, 0, st>>>(src, rest, scale, dst);
    CHZ_CUDA(cudaGetLastError());
  }
  return CHZ_OK;
}

int chz_fft_rows_dev(chz_t* h, const chz_cf32* u_dev, chz_cf32* y_dev, uint64_t nrows) {
  if (!h || ((!u_dev || !y_dev) && nrows)) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  return launch_fft_rows(h, (const float2*)u_dev, (float2*)y_dev, (long long)nrows, h->stream);
}

int chz_retained(const chz_t* h, const chz_cf32** y_dev, uint64_t* nrows) {
  if (!h) return CHZ_EINVAL;
  if (y_dev) *y_dev = (const chz_cf32*)h->d_store;
  if (nrows) *nrows = h->store_rows;
  return CHZ_OK;
}

int chz_reserve_rows(chz_t* h, uint64_t nrows) {
  if (!h) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  return ensure_store(h, nrows);
}

int chz_pdws_dev(chz_t* h, const chz_pdw_params_t* params, const chz_cf32* y_dev, uint64_t nrows,
                 chz_pdw_t* out, uint64_t cap, uint64_t* n) {
  if (!h || !params || (!y_dev && nrows)) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  const int rc = pdw_extract(h, params, (const float2*)y_dev, nrows);
  if (rc) return rc;
  if (n) *n = h->pdws.size();
  if (h->pdws.size() > cap) return CHZ_ECAPACITY;
  if (out && !h->pdws.empty()) memcpy(out, h->pdws.data(), h->pdws.size() * sizeof(chz_pdw_t));
  return CHZ_OK;
}

int chz_pdws(chz_t* h, const chz_pdw_params_t* params, chz_pdw_t* out, uint64_t cap, uint64_t* n) {
  if (!h) return CHZ_EINVAL;
  if (!h->retain) return CHZ_ESTATE;
  return chz_pdws_dev(h, params, (const chz_cf32*)h->d_store, h->store_rows, out, cap, n);
}

int chz_pdws_fetch(const chz_t* h, chz_pdw_t* out, uint64_t cap, uint64_t* n) {
  if (!h) return CHZ_EINVAL;
  if (n) *n = h->pdws.size();
  if (h->pdws.size() > cap) return CHZ_ECAPACITY;
  if (out && !h->pdws.empty()) memcpy(out, h->pdws.data(), h->pdws.size() * sizeof(chz_pdw_t));
  return CHZ_OK;
}

int chz_pdw_noise_floor(const chz_t* h, double* nf, uint32_t cap) {
  if (!h || !nf) return CHZ_EINVAL;
  if (h->noise_floor.size() != h->M) return CHZ_ESTATE;
  if (cap < h->M) return CHZ_ECAPACITY;
  memcpy(nf, h->noise_floor.data(), sizeof(double) * h->M);
  return CHZ_OK;
}

void* chz_alloc_host(uint64_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void chz_free_host(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"

