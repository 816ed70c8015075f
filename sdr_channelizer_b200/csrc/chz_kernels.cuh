// Channelizer kernels (K1 unpack, K2 polyphase FIR, K3 FFT, and the fused K1+K2+K3), sm_100a.
// Reference math replaced: matlab/create_pdws_channelized.m:35-38 (normalise) and :57
// (iq = channelizer(iq), MathWorks dsp.Channelizer — closed source; definition in DESIGN.md).
#pragma once
#ifndef CHZ_FUSED_MINB
#define CHZ_FUSED_MINB 2
#endif
#include <type_traits>

#include "chz_device.cuh"

namespace chzi {

// Everything a launch needs to know about where stream sample `idx` lives and which rows to make.
// Stream indices count complex samples since the last reset.  Row m's newest sample is m*D and it is
// produced once the whole frame [m*D, m*D + D) has arrived.
struct ChanParams {
  const void* in;        // raw samples of this call: stream indices [in_base, in_base + n_in)
  const void* hist;      // raw history: stream indices [hist_base, in_base)
  long long in_base, n_in, hist_base;
  const float* taps;     // [P][M], already multiplied by 2^-(bit_width-1) (exact)
  const float2* tw;      // W_M^i = e^{+j 2 pi i / M}, i < M
  float2* out;           // row `row_base` starts at out[0]; row-major [row][M]
  long long row_base;    // first stream row produced by this call
  long long nrows;       // rows produced by this call
  int M, D, os;          // os = M / D (1 or 2)
  int bpb;               // k_fir with run-time M: branches per 128-thread block (a divisor of M, <= 128)
  int span_rows;         // rows of one phase handled between window warm-ups (multiple of P)
  long long spans_per_phase;
};

// Raw word of stream sample idx; 0 (which unpacks to 0+0j) outside what has been received:
// x[n < 0] = 0, and rows past the end of a partial tile are computed but never stored.
template <bool IN16>
__device__ __forceinline__ uint32_t load_raw(const ChanParams& p, long long idx) {
  typedef typename RawT<IN16>::type raw_t;
  if (idx >= p.in_base) {
    if (idx < p.in_base + p.n_in) return __ldg((const raw_t*)p.in + (idx - p.in_base));
    return 0u;
  }
  if (idx >= p.hist_base) return __ldg((const raw_t*)p.hist + (idx - p.hist_base));
  return 0u;
}

// ---- K1 stand-alone: raw -> complex fp32, bit exact (create_pdws_channelized.m:35-38) ------------
// Two samples per thread: one 8-byte (int16 pairs) or 4-byte (int8 pairs) load, one 16-byte store, consecutive lanes on
// consecutive sample pairs, up to 32 blocks per SM.  Measured on 614.4 M samples (tools/ubench/mixbw.cu): 6.17 TB/s
// against 5.64 TB/s for one sample per thread with 16 blocks per SM.  `in` and `out` must be 8- / 16-byte aligned
// (the launcher checks and otherwise takes k_unpack1); n2 = number of sample PAIRS.
template <bool IN16>
__global__ void __launch_bounds__(256) k_unpack(const void* __restrict__ in, long long n2, float scale, float4* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    uint32_t w0, w1;
    if (IN16) { const uint2 q = __ldg((const uint2*)in + i); w0 = q.x; w1 = q.y; }
    else { const uint32_t q = __ldg((const uint32_t*)in + i); w0 = q & 0xffffu; w1 = q >> 16; }
    const float2 a = unpack_raw<IN16>(w0), b = unpack_raw<IN16>(w1);
    out[i] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);   // power-of-two scale: exact
  }
}
// one sample per thread: odd sample counts' last sample and misaligned buffers
template <bool IN16>
__global__ void k_unpack1(const void* __restrict__ in, long long n, float scale, float2* __restrict__ out) {
  typedef typename RawT<IN16>::type raw_t;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 v = unpack_raw<IN16>(__ldg((const raw_t*)in + i));
    out[i] = make_float2(v.x * scale, v.y * scale);
  }
}

// Describes the span a thread group works on.
struct Span {
  long long m0;      // first stream row of the span
  long long count;   // rows in the span (stride `os` rows apart)
  int skip;          // leading rows that are computed but not stored (0 or 1, see make_span)
  int shift;         // (m*D) mod M, the circular branch rotation of these rows (0 or M/2)
};
// Rows are filtered in pairs whose two FMA chains visit the taps in different (rotated) orders, so a
// row's rounding depends on whether it is the first or second of its pair.  To keep results
// independent of how a recording is cut into calls/shards, pairs are aligned to the GLOBAL row
// index: when the first row of a phase has an odd global index, the first span starts one row
// earlier and that extra row is computed but not stored.
__device__ __forceinline__ Span make_span(const ChanParams& p, long long sp) {
  Span s;
  const int phase_i = (int)(sp % p.os);          // which residue class of rows (relative to row_base)
  const long long si = sp / p.os;
  const long long first = p.row_base + phase_i;   // first row of this class
  const int lead = (int)((first / p.os) & 1);     // parity of its global per-phase index
  const long long cnt_phase = (p.nrows - phase_i + p.os - 1) / p.os + lead;
  const long long i0 = si * (long long)p.span_rows;
  s.m0 = first + (i0 - lead) * p.os;
  s.count = cnt_phase - i0;
  if (s.count > p.span_rows) s.count = p.span_rows;
  if (s.count < 0) s.count = 0;
  s.skip = si == 0 ? lead : 0;
  if (cnt_phase - lead <= 0) s.count = 0;         // no real row in this phase
  s.shift = (int)((s.m0 * p.D) % p.M);
  return s;
}

// ---- K2: polyphase FIR commutator, one thread per branch, sliding window in registers --------------
// u_p[m] = sum_q h[qM+p] x[mD - qM - p].  The P taps of a branch and its last P samples live in
// registers; a new row costs one 4-byte (int16) or 2-byte (int8) coalesced load and P packed FMAs
// (fma.rn.f32x2 on (re,im) with the tap duplicated).  `emit(i, value)` receives row i of the span.
// MT: number of channels when known at compile time (fused kernel), 0 = take prm.M.
// PF: prefetch policy for the next tile's raw samples: 0 = late (at the last rows; the fused kernel's FFT hides
// the latency), 1 = a whole tile ahead into a second buffer (FIR-only kernel), 2 = half a tile ahead, in place.
template <int P, bool IN16, int MT, int PF, typename Emit>
__device__ __forceinline__ void fir_span(const ChanParams& prm, const Span& sp, int p, Emit emit) {
  typedef typename RawT<IN16>::type raw_t;
  const long long Ml = MT ? MT : prm.M;
  float h[P];
  float2 w[P];
  #pragma unroll
  for (int q = 0; q < P; q++) h[q] = __ldg(prm.taps + q * Ml + p);
  const long long base = sp.m0 * prm.D - p;   // newest sample of span row 0 for this branch
  const long long in_end = prm.in_base + prm.n_in;
  const raw_t* __restrict__ inp = (const raw_t*)prm.in - prm.in_base;   // inp[idx] for idx in [in_base, in_end)
  // P raw words at lo, lo+M, ...: unconditional coalesced loads when the whole tile lies inside this
  // call's input (every tile but the first/last few of a call), guarded loads otherwise.
  auto load_part = [&](long long lo, uint32_t (&raw)[P], auto first, auto count) {   // raw[first + k] = x[lo + (first + k) M]
    constexpr int F = decltype(first)::value, N = decltype(count)::value;
    if (lo + F * Ml >= prm.in_base && lo + (F + N - 1) * Ml < in_end) {
      const raw_t* __restrict__ src = inp + lo;
      #pragma unroll
      for (int ii = F; ii < F + N; ii++) raw[ii] = __ldg(src + ii * Ml);
    } else {
      #pragma unroll
      for (int ii = F; ii < F + N; ii++) raw[ii] = load_raw<IN16>(prm, lo + ii * Ml);
    }
  };
  auto load_tile = [&](long long lo, uint32_t (&raw)[P]) {
    load_part(lo, raw, std::integral_constant<int, 0>{}, std::integral_constant<int, P>{});
  };
  // warm-up rows -P..-1 and tile 0 are requested back to back so their latencies overlap
  uint32_t raw[P];
  {
    uint32_t wr[P];
    load_tile(base - P * Ml, wr);
    load_tile(base, raw);
    #pragma unroll
    for (int k = 1; k < P; k++) w[P - k] = unpack_raw<IN16>(wr[P - k]);   // row -k sits in slot P-k
    w[0] = make_float2(0.f, 0.f);
  }
  // One tile = P rows filtered two at a time: two independent P-long FMA chains interleave (ILP 2)
  // with no extra adds.  Row ii uses slot (ii - q) mod P for tap q.  LATE: re-fill cur[] with the next
  // tile's samples as soon as its last word is unpacked.
  auto rows = [&](long long i0, uint32_t (&cur)[P], auto late) {
    constexpr bool LATE = decltype(late)::value;
    #pragma unroll
    for (int ii = 0; ii < P; ii += 2) {
      w[ii] = unpack_raw<IN16>(cur[ii]);
      // tap P-1 of row ii reads the OLDEST sample, which sits in the slot row ii+1 is about to take:
      // start row ii's chain with it before that slot is overwritten
      float2 a0 = __fmul2_rn(make_float2(h[P - 1], h[P - 1]), w[(ii + 1) % P]);
      if (ii + 1 < P) w[ii + 1] = unpack_raw<IN16>(cur[ii + 1]);
      // fused kernel: cur[] is fully consumed at the last rows; request the NEXT tile's samples into it
      // now, before these rows' FMAs and the FFT the caller runs in emit(): the DRAM latency hides there.
      if (LATE && ii + 2 >= P && i0 + P < sp.count) load_tile(base + (i0 + P) * Ml, cur);
      if (PF == 2 && i0 + P < sp.count) {   // HALF: refill each half of cur[] as soon as it has been consumed
        if (ii + 2 == P / 2) load_part(base + (i0 + P) * Ml, cur, std::integral_constant<int, 0>{}, std::integral_constant<int, P / 2>{});
        if (ii + 2 >= P) load_part(base + (i0 + P) * Ml, cur, std::integral_constant<int, P / 2>{}, std::integral_constant<int, P - P / 2>{});
      }
      // Row ii+1 walks the same window slots one tap later (q+1), so each step's two FMAs share their
      // 64-bit window operand (register reuse); its tap order is therefore rotated by one relative to
      // row ii -- pairs are aligned to global row parity (make_span) to keep that deterministic.
      float2 a1 = ii + 1 < P ? __fmul2_rn(make_float2(h[0], h[0]), w[ii + 1]) : make_float2(0.f, 0.f);
      #pragma unroll
      for (int q = 0; q < P - 1; q++) {
        a0 = __ffma2_rn(make_float2(h[q], h[q]), w[(ii - q + P) % P], a0);
        if (ii + 1 < P && q + 1 < P) a1 = __ffma2_rn(make_float2(h[q + 1], h[q + 1]), w[(ii - q + P) % P], a1);
      }
      emit((int)ii, i0 + ii, a0);
      if (ii + 1 < P) emit((int)ii + 1, i0 + ii + 1, a1);
    }
  };
  if constexpr (PF == 1) {
    // FIR-only kernel: nothing but P rows of FMAs separates two tiles, so the next tile is requested a
    // whole tile ahead into a second buffer (ping-pong)
    uint32_t rb[P];
    for (long long i0 = 0; i0 < sp.count; i0 += 2 * P) {
      if (i0 + P < sp.count) load_tile(base + (i0 + P) * Ml, rb);
      rows(i0, raw, std::false_type{});
      if (i0 + P < sp.count) {
        if (i0 + 2 * P < sp.count) load_tile(base + (i0 + 2 * P) * Ml, raw);
        rows(i0 + P, rb, std::false_type{});
      }
    }
  } else if constexpr (PF == 2) {
    for (long long i0 = 0; i0 < sp.count; i0 += P) rows(i0, raw, std::false_type{});
  } else {
    for (long long i0 = 0; i0 < sp.count; i0 += P) rows(i0, raw, std::true_type{});
  }
}

// Split path, kernel A: FIR only, u rows to global memory (already circularly rotated).
// grid.x = branch blocks * span blocks; block = 128 threads.
// MT = M when instantiated for a fixed channel count (immediate load/store offsets), 0 = any M.
template <int P, bool IN16, int MT>
__global__ void __launch_bounds__(128, 4) k_fir(ChanParams prm, float2* __restrict__ u) {
  const int Mv = MT ? MT : prm.M;
  // branches per block: 128, or for channel counts that are not multiples of 128 the largest divisor of M
  // the host found (560 = 5 x 112, 200 = 2 x 100, ...)
  const int bpb = MT ? (MT < 128 ? MT : 128) : prm.bpb;
  const int nbb = Mv / bpb;                        // branch blocks
  const int groups = 128 / bpb;                    // spans handled side by side in one block
  const int bb = blockIdx.x % nbb, g = threadIdx.x / bpb;
  if (g >= groups) return;                         // 128 is not a multiple of bpb (M = 56: 16 spare threads)
  // Branch p reads x[mD - p]: with lanes on branches 32w .. 32w+31 a warp's 128 bytes start 4 bytes past a line boundary
  // (5 sectors per load, and the L1's sector promotion on top: at configs[3] size k_fir reads 1.67x the recording from
  // DRAM).  Moving every thread one branch up, (p + 1) mod M, line-aligns the loads but misaligns the 8-byte stores of
  // u instead: measured 143 against 185 GS/s on configs[3] -- rejected.  Storing the row ROTATED by one element as
  // well (aligned loads, aligned stores) and reading it back one element later in the FFT kernel moves the misalignment
  // to the 8-byte reads of a kernel that already runs at 92 % of the DRAM peak: 186.9 against 210.1 GS/s at 280 M samples,
  // 182.5 against 186.4 at full size -- rejected too.
  const int p = bb * bpb + threadIdx.x % bpb;
  const long long nspans = prm.spans_per_phase * prm.os;
  const long long sstride = (long long)(gridDim.x / nbb) * groups;
  for (long long s = (long long)(blockIdx.x / nbb) * groups + g; s < nspans; s += sstride) {
    const Span sp = make_span(prm, s);
    if (sp.count <= 0) continue;
    const int r = (p - sp.shift + Mv) % Mv;        // u'[r] = u[(r + shift) mod M]
    float2* dst = u + (sp.m0 - prm.row_base) * (long long)Mv + r;
    const long long rstride = (long long)prm.os * Mv;
    fir_span<P, IN16, MT, 1>(prm, sp, p, [&](int, long long i, float2 v) {
      if (i >= sp.skip && i < sp.count) dst[i * rstride] = v;
    });
  }
}

// Split path, kernel A for tap counts without a register-window instantiation: direct evaluation,
// P loads per output (slow; correctness fallback only).
template <bool IN16>
__global__ void k_fir_any(ChanParams prm, int P, float2* __restrict__ u) {
  const long long total = prm.nrows * prm.M;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long row = e / prm.M;
    const int p = (int)(e - row * prm.M);
    const long long m = prm.row_base + row, t = m * prm.D;
    float2 acc = make_float2(0.f, 0.f);
    for (int q = 0; q < P; q++) {
      const float h = __ldg(prm.taps + q * prm.M + p);
      acc = __ffma2_rn(make_float2(h, h), unpack_raw<IN16>(load_raw<IN16>(prm, t - (long long)q * prm.M - p)), acc);
    }
    const int shift = (int)(t % prm.M);
    u[row * prm.M + (p - shift + prm.M) % prm.M] = acc;
  }
}

// Any-M row DFT (the reference's natural channel counts are not powers of two: M = fs*1e-6 = 56,
// matlab/create_pdws_channelized.m:31).  Direct evaluation y_k = sum_p u_p W_M^{kp}, O(M^2) per row,
// table W_M^i in shared memory indexed by (k p) mod M.  Functional path, not tuned.
static __global__ void __launch_bounds__(256) k_dft_rows_any(const float2* u, float2* y, const float2* __restrict__ tw_g, int M,
                                                      long long nrows) {
  extern __shared__ float2 smem[];
  float2* row = smem;
  float2* tw = smem + M;
  for (int i = threadIdx.x; i < M; i += blockDim.x) tw[i] = tw_g[i];
  for (long long r = blockIdx.x; r < nrows; r += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) row[i] = u[r * M + i];
    __syncthreads();
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
      float2 acc = make_float2(0.f, 0.f);
      int idx = 0;
      for (int p = 0; p < M; p++) {
        const float2 a = row[p], w = tw[idx];
        acc.x = fmaf(a.x, w.x, fmaf(-a.y, w.y, acc.x));
        acc.y = fmaf(a.x, w.y, fmaf(a.y, w.x, acc.y));
        idx += k;
        if (idx >= M) idx -= M;
      }
      y[r * M + k] = acc;
    }
  }
}

// ---- any 7-smooth M: run-time mixed-radix Stockham row FFT ------------------------------------------
// Channel counts follow the radio's sample rate (M = fs*1e-6, matlab/create_pdws_channelized.m:31): 40,
// 48, 80, 100, 112, 120, 200 ... are as natural as 56.  Only 56 and 560 have compile-time plans; every
// other M whose prime factors are 2, 3, 5, 7 runs this kernel: radices from {16, 8, 4, 2, 3, 5, 7} chosen on
// the host (MixedPlan), one Stockham pass per radix in shared memory (unpadded), twiddles from the same
// per-pass table layout as the compiled plans.  M with a larger prime factor keeps the O(M^2) DFT below.
struct MixedPlan { int np; int r[12]; };

template <int R>
__device__ __forceinline__ void mixed_pass(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ tw,
                                           int M, int NS, int rows, int t, int NT, float2* __restrict__ gout, int vrows) {
  const int BPR = M / R, total = rows * BPR;
  for (int b = t; b < total; b += NT) {
    const int row = b / BPR, j = b - row * BPR;
    const float2* s = src + row * M;
    float2 v[R];
    #pragma unroll
    for (int q = 0; q < R; q++) v[q] = s[j + q * BPR];
    const int k = j % NS;
    if (NS > 1) {
      #pragma unroll
      for (int q = 1; q < R; q++) v[q] = cmul(v[q], tw[(q - 1) * NS + k]);     // W_{NS R}^{q k}
    }
    dft<R>(v);
    const int j0 = (j - k) * R + k;
    if (gout) {
      if (row < vrows) {
        float2* g = gout + (long long)row * M;
        #pragma unroll
        for (int q = 0; q < R; q++) g[j0 + q * NS] = v[q];
      }
    } else {
      float2* d = dst + row * M;
      #pragma unroll
      for (int q = 0; q < R; q++) d[j0 + q * NS] = v[q];
    }
  }
}

// One block transforms `rows_per_block` rows at a time.  Dynamic smem: 2 * rows_per_block * M float2.
static __global__ void __launch_bounds__(256) k_fft_rows_mixed(const float2* u, float2* y, const float2* __restrict__ tw_g, int M,
                                                               long long nrows, int rows_per_block, MixedPlan plan) {
  extern __shared__ float2 smem[];
  float2* buf0 = smem;
  float2* buf1 = smem + (size_t)rows_per_block * M;
  const int t = threadIdx.x;
  for (long long r0 = (long long)blockIdx.x * rows_per_block; r0 < nrows; r0 += (long long)gridDim.x * rows_per_block) {
    const int vrows = (int)((nrows - r0) < rows_per_block ? (nrows - r0) : rows_per_block);
    __syncthreads();
    for (int e = t; e < vrows * M; e += 256) buf0[e] = u[r0 * M + e];
    __syncthreads();
    float2* src = buf0;
    float2* dst = buf1;
    int NS = 1, off = 0;
    for (int pass = 0; pass < plan.np; pass++) {
      const int R = plan.r[pass];
      float2* gout = pass == plan.np - 1 ? y + r0 * M : nullptr;
      const float2* tw = tw_g + off;
      switch (R) {
        case 16: mixed_pass<16>(src, dst, tw, M, NS, vrows, t, 256, gout, vrows); break;
        case 8: mixed_pass<8>(src, dst, tw, M, NS, vrows, t, 256, gout, vrows); break;
        case 7: mixed_pass<7>(src, dst, tw, M, NS, vrows, t, 256, gout, vrows); break;
        case 5: mixed_pass<5>(src, dst, tw, M, NS, vrows, t, 256, gout, vrows); break;
        case 4: mixed_pass<4>(src, dst, tw, M, NS, vrows, t, 256, gout, vrows); break;
        case 3: mixed_pass<3>(src, dst, tw, M, NS, vrows, t, 256, gout, vrows); break;
        default: mixed_pass<2>(src, dst, tw, M, NS, vrows, t, 256, gout, vrows); break;
      }
      if (NS > 1) off += (R - 1) * NS;                   // the first pass has no twiddles and no table entries
      NS *= R;
      float2* sw = src; src = dst; dst = sw;
      __syncthreads();
    }
  }
}

// Split path, kernel B (also K3 on its own): M-point FFT of rows in global memory.
// One block transforms ROWS rows at a time.  Dynamic smem: 2 * ROWS * RowStride<M> float2 + M float2.
template <int M, int ROWS, int NT>
__global__ void __launch_bounds__(NT) k_fft_rows(const float2* u, float2* y,   // launched in place (u == y): no __restrict__
                                                
                                                 const float2* __restrict__ tw_g, long long nrows) {
  extern __shared__ float2 smem[];
  constexpr int S = RowStride<M>::value;
  float2* buf0 = smem;
  float2* buf1 = buf0 + ROWS * S;
  float2* tw = buf1 + ROWS * S;
  for (int i = threadIdx.x; i < M; i += NT) tw[i] = tw_g[i];
  for (long long r0 = (long long)blockIdx.x * ROWS; r0 < nrows; r0 += (long long)gridDim.x * ROWS) {
    const int vrows = (int)((nrows - r0) < ROWS ? (nrows - r0) : ROWS);
    __syncthreads();
    for (int e = threadIdx.x; e < ROWS * M; e += NT) {
      const int row = e / M, i = e - row * M;
      buf0[row * S + padi_first<M>(i)] = row < vrows ? u[(r0 + row) * M + i] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft_tile_to_global<M, ROWS, NT, false>(buf0, buf1, tw, nullptr, threadIdx.x, y + r0 * M, (long long)M, 0, vrows,
                                           [] { __syncthreads(); });
  }
}

// Row FFT for the large sizes of the split path (M = 512..4096, first radix 16): ROWS*M = 4096
// elements per block iteration, 256 threads, exactly one radix-16 butterfly per thread in the first
// pass.  That pass reads its 16 operands straight from global memory into registers (no staging
// copy), and the operands of the block's NEXT rows are requested before this iteration's passes run,
// so DRAM latency overlaps the shared-memory passes.  Twiddles come from the global table through L1.
// In-place safe (a block only reads and writes its own rows).
template <int M, int ROWS>
__global__ void __launch_bounds__(256, 2) k_fft_rows_big(const float2* u, float2* y, const float2* __restrict__ tw_g,
                                                         long long nrows) {
  typedef Plan<M> PL;
  static_assert(PL::np == 3 && PL::r0 == 16 && ROWS * M == 4096, "large-M plan expected");
  extern __shared__ float2 smem[];
  constexpr int S = RowStride<M>::value, BPR0 = M / 16;
  float2* bufA = smem;
  float2* bufB = bufA + ROWS * S;
  const int t = threadIdx.x, row = t / BPR0, j = t % BPR0;
  auto load = [&](long long r0, float2 (&v)[16]) {
    const bool ok = r0 + row < nrows;
    const float2* src = u + (r0 + row) * (long long)M + j;
    #pragma unroll
    for (int q = 0; q < 16; q++) v[q] = ok ? src[q * BPR0] : make_float2(0.f, 0.f);
  };
  const long long step = (long long)gridDim.x * ROWS;
  float2 cur[16];
  long long r0 = (long long)blockIdx.x * ROWS;
  if (r0 < nrows) load(r0, cur);
  for (; r0 < nrows; r0 += step) {
    float2 nxt[16];
    if (r0 + step < nrows) load(r0 + step, nxt);
    const long long left = nrows - r0;
    const int vhi = (int)(left < ROWS ? left : ROWS);
    // pass 1 (radix 16, no twiddles) from registers, Stockham-transposed into bufA
    dft<16>(cur);
    {
      float2* d = bufA + row * S;
      #pragma unroll
      for (int q = 0; q < 16; q++) d[padi<M>(j * 16 + q)] = cur[q];
    }
    __syncthreads();
    stockham_pass<M, PL::r1, PL::r0, ROWS, 256, false, false>(bufA, bufB, tw_g, nullptr, t, nullptr, 0, 0, 0);
    __syncthreads();
    stockham_pass<M, PL::r2, PL::r0 * PL::r1, ROWS, 256, true, false>(bufB, bufA, tw_g, nullptr, t, y + r0 * (long long)M,
                                                                      (long long)M, 0, vhi);
    // no barrier needed here: the next iteration writes bufA (last read before the barrier above) and
    // only touches bufB after its own first barrier, which every thread reaches after finishing this pass
    #pragma unroll
    for (int q = 0; q < 16; q++) cur[q] = nxt[q];
  }
}

// ---- fused K1+K2+K3: one global read (raw int samples), one global write (fp32 channels) ----------
// A block is G groups of M threads; a group walks spans of rows.  Every RT filtered rows (RT divides
// P) the group runs the M-point FFT on its shared tile and streams the result to global memory.
template <int M, int P> struct FusedCfg {
  static constexpr int MG = GroupThreads<M>::value;  // threads per group (M rounded up to whole warps)
  static constexpr int NT = MG < 256 ? 256 / MG * MG : MG;   // threads per block
  static constexpr int G = NT / MG;                  // groups per block
  // rows per FFT tile: largest divisor of P keeping the two tile buffers of a block under ~72 KB (several
  // blocks per SM); blocks of >= 512 threads are alone on their SM anyway (registers) and take up to 200 KB:
  // twice the rows per barrier round (M = 512: barrier stalls were 23 % of all samples with 8-row tiles)
  static constexpr int RTMAX = ((NT >= 512 ? 200 : 72) * 1024) / (2 * 8 * RowStride<M>::value * G);
  static constexpr int RT = RTMAX >= P ? P : (P % 8 == 0 && RTMAX >= 8 ? 8 : (P % 6 == 0 && RTMAX >= 6 ? 6 : (P % 4 == 0 && RTMAX >= 4 ? 4 : (P % 2 == 0 && RTMAX >= 2 ? 2 : 1))));
  // float2 per group (two tile buffers).  M = 8: four groups share a warp and two a half-warp (the unit a
  // 64-bit shared access is processed in); 8 elements of padding put neighbouring groups 16 banks apart
  // (measured without it: half of all shared wavefronts were bank conflicts, LSU data pipe 98 % busy).
  static constexpr int GSTRIDE = 2 * RT * RowStride<M>::value + (M == 8 ? 8 : 0);
  static constexpr size_t SMEM = (size_t)(M + G * GSTRIDE) * sizeof(float2);
};

template <int M> __device__ __forceinline__ void group_sync(int g) {
  if (M > 32) asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GroupThreads<M>::value) : "memory");   // named barrier per group
  else __syncwarp();   // M <= 32: a group lives inside one warp
}

template <int M, int P, bool IN16>
__global__ void __launch_bounds__(FusedCfg<M, P>::NT, FusedCfg<M, P>::NT <= 256 ? CHZ_FUSED_MINB : 1) k_chan_fused(ChanParams prm) {
  extern __shared__ float2 smem[];
  typedef FusedCfg<M, P> CF;
  constexpr int NT = CF::NT, G = CF::G, RT = CF::RT, S = RowStride<M>::value, MG = CF::MG;
  constexpr bool TWREG = TwReg<M, MG>::value;
  float2* tw = smem;                                  // M twiddles
  const int g = threadIdx.x / MG, t = threadIdx.x % MG;
  // non-power-of-two M: the spare lanes of the last warp shadow branch M-1 through the FIR (same control
  // flow, so they meet every barrier) without storing, and take their share of the FFT butterflies
  const bool branch = MG == M || t < M;
  // (thread t on branch (t + 1) mod M would make a warp's 32 samples x[mD - p] one aligned line; measured: no gain at
  // M = 64 (435.6 against 441.0 GS/s), a loss at M = 128 -- the loads' extra sector is not what bounds this kernel)
  const int p = branch ? t : M - 1;
  float2* buf0 = smem + M + (size_t)g * CF::GSTRIDE;   // [RT][S]
  float2* buf1 = buf0 + RT * S;
  for (int i = threadIdx.x; i < M; i += NT) tw[i] = prm.tw[i];
  float2 twr[TwReg<M, MG>::count];
  load_last_pass_twiddles<M, MG>(prm.tw, t, twr);
  __syncthreads();
  const long long nspans = prm.spans_per_phase * prm.os;
  const long long rstride = (long long)prm.os * M;
  for (long long s = (long long)blockIdx.x * G + g; s < nspans; s += (long long)gridDim.x * G) {
    const Span sp = make_span(prm, s);
    if (sp.count <= 0) continue;                      // the whole group takes the same branch
    const int r = padi_first<M>((p - sp.shift + M) % M);
    float2* gout = prm.out + (sp.m0 - prm.row_base) * (long long)M;
    fir_span<P, IN16, M, 0>(prm, sp, p, [&](int ii, long long i, float2 v) {
      if (branch) buf0[(ii % RT) * S + r] = v;
      if (ii % RT == RT - 1) {
        group_sync<M>(g);
        const long long i0 = i - (RT - 1);
        const long long left = sp.count - i0;
        const int vhi = (int)(left < RT ? (left < 0 ? 0 : left) : RT);       // rows [vlo, vhi) of the tile are stored
        const int vlo = i0 < sp.skip ? (int)(sp.skip - i0) : 0;
        fft_tile_to_global<M, RT, MG, TWREG>(buf0, buf1, tw, twr, t, gout + i0 * rstride, rstride, vlo, vhi,
                                             [&] { group_sync<M>(g); });
        if (Plan<M>::np != 2) group_sync<M>(g);   // the last pass of 1- and 3-pass plans reads buf0
      }
    });
  }
}

}  // namespace chzi
