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
template <bool IN16>
__global__ void k_unpack(const void* __restrict__ in, long long n, float scale, float2* __restrict__ out) {
  typedef typename RawT<IN16>::type raw_t;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 v = unpack_raw<IN16>(__ldg((const raw_t*)in + i));
    out[i] = make_float2(v.x * scale, v.y * scale);   // power-of-two scale: exact
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
__global__ void __launch_bounds__(NT) k_fft_rows(const float2* __restrict__ u, float2* __restrict__ y,
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
      buf0[row * S + padi<M>(i)] = row < vrows ? u[(r0 + row) * M + i] : make_float2(0.f, 0.f);
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
    const int r = padi<M>((p - sp.shift + M) % M);
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

// ---- warp-specialised fused kernel (M = 64): FIR warps and FFT warps ------------------------------
// The plain fused kernel alternates two phases in every warp: the FIR (FMA-pipe bound) and the FFT
// (shared-memory / latency bound), with only 16 warps per SM because every thread carries the FIR
// state AND the FFT registers.  Here a block is two warpgroups: warps 0-3 run the FIR of two groups
// (64 branches each) and nothing else, warps 4-7 run the FFT of those two groups.  The FIR warpgroup
// raises its register budget (setmaxnreg.inc), the FFT warpgroup lowers it (setmaxnreg.dec), so three
// blocks (24 warps) fit an SM and the two kinds of work overlap instead of alternating.
// Hand-off: per group a ring of WS_NB tile buffers in shared memory guarded by mbarriers
// (full[b]: 64 FIR threads arrive after writing a tile; empty[b]: 64 FFT threads arrive after their
// first pass has consumed it).
constexpr int WS_NB = 2;
constexpr int WS_FIR_REGS = 104, WS_FFT_REGS = 56;

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

template <int M, int P> struct WsCfg {
  static constexpr int S = RowStride<M>::value;
  static constexpr int TILE = P * S;                                   // float2 per tile buffer
  // per group: WS_NB ring buffers + one FFT scratch buffer; per block: 2 groups + twiddles + barriers
  static constexpr size_t SMEM = (size_t)(2 * (WS_NB + 1) * TILE + M) * sizeof(float2) + 2 * 2 * WS_NB * sizeof(uint64_t);
};

template <int M, int P, bool IN16>
__global__ void __launch_bounds__(256, 3) k_chan_ws(ChanParams prm) {
  static_assert(M == 64, "two warps per group");
  typedef WsCfg<M, P> WC;
  constexpr int S = WC::S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw = (float2*)smem_raw;                                      // M twiddles
  float2* bufs = tw + M;                                               // [2 groups][WS_NB + 1][P][S]
  uint64_t* bars = (uint64_t*)(bufs + 2 * (WS_NB + 1) * WC::TILE);     // [2 groups][full WS_NB | empty WS_NB]
  const int warp = threadIdx.x >> 5;
  const bool is_fir = warp < 4;
  const int g = (warp & 3) >> 1;                                       // group inside the block
  const int p = threadIdx.x & 63;                                      // branch (FIR) / FFT thread index in the group
  float2* ring = bufs + (size_t)g * (WS_NB + 1) * WC::TILE;
  float2* scratch = ring + WS_NB * WC::TILE;
  uint64_t* full = bars + g * 2 * WS_NB;
  uint64_t* empty = full + WS_NB;
  for (int i = threadIdx.x; i < M; i += 256) tw[i] = prm.tw[i];
  if (threadIdx.x < 2 * 2 * WS_NB) mbar_init(bars + threadIdx.x, 64);
  __syncthreads();
  const long long nspans = prm.spans_per_phase * prm.os;
  const long long rstride = (long long)prm.os * M;
  const long long gg = (long long)blockIdx.x * 2 + g, gstride = (long long)gridDim.x * 2;
  unsigned tile = 0;                                                   // tiles handled so far by this group
  if (is_fir) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_FIR_REGS));
    for (long long s = gg; s < nspans; s += gstride) {
      const Span sp = make_span(prm, s);
      if (sp.count <= 0) continue;
      const int r = padi<M>((p - sp.shift + M) % M);
      fir_span<P, IN16, M, 2>(prm, sp, p, [&](int ii, long long, float2 v) {
        const unsigned b = tile % WS_NB, n = tile / WS_NB;
        if (ii == 0 && n > 0) mbar_wait(&empty[b], (n - 1) & 1);       // the FFT warps are done with this buffer
        ring[b * WC::TILE + ii * S + r] = v;
        if (ii == P - 1) { mbar_arrive(&full[b]); tile++; }
      });
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_FFT_REGS));
    for (long long s = gg; s < nspans; s += gstride) {
      const Span sp = make_span(prm, s);
      if (sp.count <= 0) continue;
      float2* gout = prm.out + (sp.m0 - prm.row_base) * (long long)M;
      for (long long i0 = 0; i0 < sp.count; i0 += P) {
        const unsigned b = tile % WS_NB, n = tile / WS_NB;
        const long long left = sp.count - i0;
        const int vhi = (int)(left < P ? left : P);
        const int vlo = i0 < sp.skip ? (int)(sp.skip - i0) : 0;
        mbar_wait(&full[b], n & 1);
        stockham_pass<M, Plan<M>::r0, 1, P, 64, false, false>(ring + b * WC::TILE, scratch, tw, nullptr, p, nullptr, 0, 0, 0);
        mbar_arrive(&empty[b]);                                        // ring buffer b may be refilled
        asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory");
        stockham_pass<M, Plan<M>::r1, Plan<M>::r0, P, 64, true, false>(scratch, nullptr, tw, nullptr, p, gout + i0 * rstride,
                                                                      rstride, vlo, vhi);
        asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory");       // scratch is free for the next tile
        tile++;
      }
    }
  }
}

// ---- large M (1024..4096): one fused launch per call on thread-block clusters ----------------------
// A branch's register window times M threads does not fit one SM, so M/512 CTAs (2, 4 or 8) form a
// cluster: CTA `rank` owns the 512 contiguous branches [512 rank, 512 rank + 512) for the FIR (same
// register-window code, coalesced 2 KB loads).  Each tile of P rows goes to a per-cluster scratch ring
// (2 tiles x P x M float2, <= 1 MB, rewritten continuously so it lives in L2 and never reaches DRAM);
// after ONE cluster barrier per tile every CTA transforms P/C whole rows of that tile (first radix-16
// pass straight from the scratch into registers, two more passes in shared memory) and streams them
// to the output.  The ring is double buffered: tile t+2 reuses tile t's slot only after barrier t+1,
// which every CTA reaches after finishing its FFT of tile t.
// DRAM traffic is the fused kernel's (raw in once, fp32 out once); the intermediate costs L2 bandwidth.
template <int M, int P, int TPC = 512> struct ClusterCfg {
  static constexpr int C = M / TPC;                  // CTAs per cluster
  static constexpr int RPC = P / C;                  // rows each CTA transforms per tile
  static constexpr bool ok = (M == 1024 || M == 2048 || M == 4096) && C >= 2 && C <= 16 && (P % C == 0) && RPC >= 1 &&
                             (TPC == 512 || RPC * (M / 16) == TPC);
  // two FFT tile buffers + the inter-pass twiddle table (the cluster barrier invalidates L1 every tile,
  // so twiddles read through L1 would come from L2 again each time)
  static constexpr size_t SMEM = ((size_t)2 * (RPC > 0 ? RPC : 1) * RowStride<M>::value + M) * sizeof(float2);
};

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// TPC = threads (= branches) per CTA.  512: one CTA per SM (all registers).  256 (CHZ_OPT_FORCE_PATH=7,8):
// twice as many CTAs per cluster, two CTAs of DIFFERENT clusters share an SM, so one cluster's barrier /
// L2 latency overlaps the other's FIR or FFT.
// PIPE (CHZ_OPT_FORCE_PATH=8): software pipelining across the cluster barrier.  After the FIR of tile t a CTA
// only ARRIVES (release) at the barrier, waits for the barrier of tile t-1 -- which everyone reached a whole
// FIR tile ago -- and transforms tile t-1, so neither the barrier round trip nor the slowest CTA of the
// cluster is on the critical path.  The ring then needs 4 slots: a CTA can be writing tile t+3 while a slow
// one still reads tile t (it is only known to have arrived for t+1).
template <int M, int P, bool IN16, int TPC, bool PIPE>
__global__ void __launch_bounds__(TPC, TPC == 512 ? 1 : 2) k_chan_cluster(ChanParams prm, float2* __restrict__ scratch) {
  typedef ClusterCfg<M, P, TPC> CC;
  typedef Plan<M> PL;
  static_assert(PL::np == 3 && PL::r0 == 16, "large-M plan expected");
  constexpr int C = CC::C, RPC = CC::RPC, S = RowStride<M>::value, BPR0 = M / 16, NSLOT = PIPE ? 4 : 2;
  extern __shared__ float2 smem[];
  float2* bufA = smem;
  float2* bufB = bufA + RPC * S;
  float2* tw = bufB + RPC * S;
  const int t = threadIdx.x;
  for (int i = t; i < M; i += TPC) tw[i] = prm.tw[i];
  __syncthreads();
  const int rank = (int)cluster_ctarank();
  const long long cid = blockIdx.x / C, ncl = gridDim.x / C;
  const int p = rank * TPC + t;
  float2* ring = scratch + (size_t)cid * NSLOT * P * M;
  const long long nspans = prm.spans_per_phase * prm.os;
  const long long rstride = (long long)prm.os * M;
  const int frow = t / BPR0, fj = t % BPR0;            // this thread's first-pass butterfly: (row, column)
  unsigned tile = 0;
  // transform rows [rank*RPC, rank*RPC + RPC) of the tile in `slot` and stream them to gout (+ row stride):
  // passes 1 and 2 (ring -> shared memory), then pass 3 (shared memory -> global)
  auto fft_tile_a = [&](const float2* slot) {
    if (frow < RPC) {                                  // pass 1 (radix 16) from the ring, bypassing L1 (another SM wrote it)
      float2 x[16];
      const float2* src = slot + (size_t)(rank * RPC + frow) * M + fj;
      #pragma unroll
      for (int q = 0; q < 16; q++) x[q] = __ldcg(src + q * BPR0);
      dft<16>(x);
      float2* d = bufA + frow * S;
      #pragma unroll
      for (int q = 0; q < 16; q++) d[padi<M>(fj * 16 + q)] = x[q];
    }
    __syncthreads();
    stockham_pass<M, PL::r1, PL::r0, RPC, TPC, false, false>(bufA, bufB, tw, nullptr, t, nullptr, 0, 0, 0);
    __syncthreads();
  };
  auto fft_tile_b = [&](float2* gout0, int vlo, int vhi) {
    stockham_pass<M, PL::r2, PL::r0 * PL::r1, RPC, TPC, true, false>(bufB, bufA, tw, nullptr, t, gout0, rstride, vlo, vhi);
  };
  // PIPE: the tile whose barrier has been arrived at but whose FFT is still to do
  bool pend = false;
  const float2* pend_slot = nullptr;
  float2* pend_gout = nullptr;
  int pend_vlo = 0, pend_vhi = 0;
  for (long long s = cid; s < nspans; s += ncl) {      // cluster-uniform loop
    const Span sp = make_span(prm, s);
    if (sp.count <= 0) continue;
    const int r = (p - sp.shift + M) % M;              // circular shift of the oversampled odd rows
    float2* gout = prm.out + (sp.m0 - prm.row_base) * (long long)M;
    fir_span<P, IN16, M, 0>(prm, sp, p, [&](int ii, long long i, float2 v) {
      float2* slot = ring + (size_t)(tile % NSLOT) * P * M;
      slot[(size_t)ii * M + r] = v;
      if (ii == P - 1) {
        const long long i0 = i - (P - 1) + rank * RPC; // first span row this CTA transforms
        const long long left = sp.count - i0;
        const int vhi = (int)(left < RPC ? (left < 0 ? 0 : left) : RPC);
        const int vlo = i0 < sp.skip ? (int)(sp.skip - i0) : 0;
        if constexpr (PIPE) {
          // The release fence of the arrive waits for every store this CTA has in flight.  Placed between
          // passes 2 and 3 of the previous tile's FFT it finds this tile's ring stores (issued a thousand
          // cycles ago) and the previous y rows (a whole tile ago) already acknowledged; right after the FIR
          // it cost a third of all stall samples (profiles/r01k).
          if (pend) { cluster_wait(); fft_tile_a(pend_slot); }   // tile-1 is complete in the ring
          cluster_arrive();                            // my part of this tile is written
          if (pend) fft_tile_b(pend_gout, pend_vlo, pend_vhi);
          pend = true; pend_slot = slot; pend_gout = gout + i0 * rstride; pend_vlo = vlo; pend_vhi = vhi;
        } else {
          cluster_barrier();                           // the whole tile is in the ring (L2)
          fft_tile_a(slot);
          fft_tile_b(gout + i0 * rstride, vlo, vhi);
        }
        tile++;
      }
    });
  }
  if constexpr (PIPE) {
    if (pend) { cluster_wait(); fft_tile_a(pend_slot); fft_tile_b(pend_gout, pend_vlo, pend_vhi); }
  }
}

// ---- large M fused over distributed shared memory: st.async + mbarrier transaction counts -------------
// Profile of the L2-ring cluster kernel (profiles/r01k): a third of all stall samples sit on the MEMBAR /
// ERRBAR of `barrier.cluster.arrive.release` (every tile each CTA must drain its ring stores to L2 before
// it may signal), the ring costs 16 B/sample of L2 bandwidth and half of it is written back to DRAM.
// Here the FIR threads send every filtered value straight into the shared memory of the CTA that will
// transform that row (`st.async.shared::cluster ... mbarrier::complete_tx::bytes`); the receiver waits on
// a local mbarrier until RPC*M*8 bytes have landed.  No fence, no ring, no global intermediate.
//   cluster = M/256 CTAs of 256 threads (= branches), two CTAs (of different clusters) per SM;
//   tile t  = P rows; CTA `rank` transforms rows [rank*RPC, rank*RPC + RPC) of every tile (RPC = P/C);
//   IN[2]   = receive buffers [RPC][RowStride] (padded layout, also the Stockham scratch of passes 2/3);
//   flow control: a CTA may send tile t only after every CTA has finished the FFT of tile t-2 (same IN slot):
//   one RELAXED cluster barrier per tile, arrive after the FFT of tile t-1, wait before the first send of t+1.
// Software pipeline per CTA: FIR(t) [sends] -> FFT(t-1) [data arrived a whole FIR ago].
template <int M, int P> struct DsmCfg {
  static constexpr int TPC = 256;
  static constexpr int C = M / TPC;
  static constexpr int RPC = P / (C > 0 ? C : 1);
  static constexpr bool ok = (M == 1024 || M == 2048 || M == 4096) && (P % C == 0) && RPC * (M / 16) == TPC;
  static constexpr size_t SMEM = (size_t)3 * (RPC > 0 ? RPC : 1) * RowStride<M>::value * sizeof(float2);
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned local, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f2(unsigned remote_addr, float2 v, unsigned remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(remote_addr), "f"(v.x), "f"(v.y), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}"
               ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_plain() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }

template <int M, int P, bool IN16>
__global__ void __launch_bounds__(256, 2) k_chan_dsm(ChanParams prm) {
  typedef DsmCfg<M, P> DC;
  typedef Plan<M> PL;
  static_assert(PL::np == 3 && PL::r0 == 16, "large-M plan expected");
  constexpr int TPC = DC::TPC, C = DC::C, RPC = DC::RPC, S = RowStride<M>::value, BPR0 = M / 16;
  constexpr unsigned TILE_BYTES = (unsigned)(RPC * M * sizeof(float2));
  extern __shared__ float2 smem[];
  __shared__ __align__(8) uint64_t full[2];
  float2* in0 = smem;                                  // IN[0] / Stockham scratch B of even tiles
  float2* in1 = in0 + RPC * S;
  float2* bufA = in1 + RPC * S;
  const int t = threadIdx.x;
  const unsigned rank = cluster_ctarank();
  if (t == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&full[0], TILE_BYTES);               // tiles 0 and 1
    mbar_expect_tx(&full[1], TILE_BYTES);
  }
  cluster_barrier();                                   // barriers initialised cluster-wide before anyone sends
  const long long cid = blockIdx.x / C, ncl = gridDim.x / C;
  const int p = (int)rank * TPC + t;
  const long long nspans = prm.spans_per_phase * prm.os;
  const long long rstride = (long long)prm.os * M;
  const int frow = t / BPR0, fj = t % BPR0;            // first-pass butterfly of this thread: (row, column)
  const unsigned in_addr[2] = {smem_u32(in0), smem_u32(in1)};
  const unsigned full_addr[2] = {smem_u32(&full[0]), smem_u32(&full[1])};
  unsigned tile = 0;                                   // tiles sent so far
  // FFT of the tile in slot s (tile index n): wait for its bytes, three passes, stream rows to gout0
  auto fft_tile = [&](unsigned n, float2* gout0, int vlo, int vhi) {
    const unsigned sl = n & 1;
    float2* in = sl ? in1 : in0;
    mbar_wait(&full[sl], (n >> 1) & 1);
    {
      float2 x[16];
      const float2* src = in + frow * S;
      #pragma unroll
      for (int q = 0; q < 16; q++) x[q] = src[padi<M>(fj + q * BPR0)];
      dft<16>(x);
      float2* d = bufA + frow * S;
      #pragma unroll
      for (int q = 0; q < 16; q++) d[padi<M>(fj * 16 + q)] = x[q];
    }
    __syncthreads();                                   // IN[sl] fully consumed: it becomes the pass-2 output
    if (t == 0) mbar_expect_tx(&full[sl], TILE_BYTES);  // arm the slot for tile n+2 (senders are held by the cluster barrier)
    stockham_pass<M, PL::r1, PL::r0, RPC, TPC, false, false>(bufA, in, prm.tw, nullptr, t, nullptr, 0, 0, 0);
    __syncthreads();
    stockham_pass<M, PL::r2, PL::r0 * PL::r1, RPC, TPC, true, false>(in, bufA, prm.tw, nullptr, t, gout0, rstride, vlo, vhi);
  };
  bool pend = false;
  unsigned pend_n = 0;
  float2* pend_gout = nullptr;
  int pend_vlo = 0, pend_vhi = 0;
  for (long long s = cid; s < nspans; s += ncl) {      // cluster-uniform loop
    const Span sp = make_span(prm, s);
    if (sp.count <= 0) continue;
    const unsigned pos = (unsigned)padi<M>((p - sp.shift + M) % M) * (unsigned)sizeof(float2);
    float2* gout = prm.out + (sp.m0 - prm.row_base) * (long long)M;
    fir_span<P, IN16, M, 0>(prm, sp, p, [&](int ii, long long i, float2 v) {
      const unsigned sl = tile & 1;
      // slot sl was last used by tile-2: every CTA has finished that FFT once the barrier it arrived at
      // after it completes (first tiles: nothing to wait for)
      if (ii == 0 && tile >= 2) cluster_wait_plain();
      const unsigned dst = (unsigned)ii / RPC;          // CTA that transforms this row
      const unsigned row_off = ((unsigned)ii % RPC) * (unsigned)(S * sizeof(float2));
      st_async_f2(mapa_u32(in_addr[sl] + row_off + pos, dst), v, mapa_u32(full_addr[sl], dst));
      if (ii == P - 1) {
        const long long i0 = i - (P - 1) + rank * RPC; // first span row this CTA transforms
        const long long left = sp.count - i0;
        const int vhi = (int)(left < RPC ? (left < 0 ? 0 : left) : RPC);
        const int vlo = i0 < sp.skip ? (int)(sp.skip - i0) : 0;
        if (pend) {
          fft_tile(pend_n, pend_gout, pend_vlo, pend_vhi);
          cluster_arrive_relaxed();                    // this CTA is done with tile pend_n's slot
        }
        pend = true; pend_n = tile; pend_gout = gout + i0 * rstride; pend_vlo = vlo; pend_vhi = vhi;
        tile++;
      }
    });
  }
  if (pend) {
    // a wait is still owed for every arrive whose matching wait was never reached (at most one)
    if (tile >= 2) cluster_wait_plain();
    fft_tile(pend_n, pend_gout, pend_vlo, pend_vhi);
  }
  cluster_barrier();                                   // nobody leaves while a peer may still send to it or wait for it
}

// ---- M = 1024 fused on CTA pairs: decimation-in-time split over distributed shared memory ----------
// 1024 branch windows do not fit one SM's registers, 512 do.  A cluster of two CTAs splits the branches
// by parity: CTA c filters branches p = 2t + c (t = thread) with the usual register windows, runs a
// 512-point FFT of its half in shared memory and leaves E = FFT512(even part) or O = FFT512(odd part)
// in a result buffer.  After ONE cluster barrier per tile the last radix-2 stage
//     Y[k] = E[k] + W_1024^k O[k],   Y[k + 512] = E[k] - W_1024^k O[k]
// is computed by both CTAs, each for 256 values of k, reading the partner's half through DSMEM
// (ld.shared::cluster) and storing two contiguous 2 KB runs per row.  Only half of the FFT output
// crosses the SM-to-SM network (4 B per output sample); DRAM traffic is the algorithmic minimum.
// The price: each CTA touches every input sector but uses half of it (8 instead of 4 B per sample
// from L2).  Result buffers are double buffered, so one barrier per tile suffices.
template <int P> struct Dit2Cfg {
  static constexpr int MS = 512;                                   // sub-FFT size
  static constexpr int RT = (P % 8 == 0) ? 8 : 4;                  // rows per FFT tile (divides P, even)
  static constexpr int S = RowStride<MS>::value;
  static constexpr size_t SMEM = ((size_t)2 * RT * S + 2 * RT * MS + MS) * sizeof(float2);
};

__device__ __forceinline__ float2 ld_dsmem(const float2* local, unsigned peer) {
  unsigned la = (unsigned)__cvta_generic_to_shared(local), ra;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(peer));
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ra));
  return v;
}

template <int P, bool IN16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) k_chan_dit2(ChanParams prm) {
  constexpr int M = 1024, MS = Dit2Cfg<P>::MS, RT = Dit2Cfg<P>::RT, S = Dit2Cfg<P>::S;
  typedef Plan<MS> PL;                                             // 512 = 16 * 8 * 4
  extern __shared__ float2 smem[];
  float2* buf0 = smem;                                             // [RT][S]
  float2* buf1 = buf0 + RT * S;
  float2* res = buf1 + RT * S;                                     // [2][RT][MS]  E or O, natural order
  float2* tw = res + 2 * RT * MS;                                  // twiddles of the 512-point plan
  const int t = threadIdx.x;
  const unsigned rank = cluster_ctarank();
  {  // inter-pass twiddles, layout (q-1)*NS + k (see stockham_pass): pass 2 (NS=16, R=8), pass 3 (NS=128, R=4)
    constexpr int N2 = (PL::r1 - 1) * PL::r0, N3 = (PL::r2 - 1) * PL::r0 * PL::r1;
    for (int i = t; i < N2 + N3; i += 512) {
      int q, k, n;
      if (i < N2) { q = i / PL::r0 + 1; k = i % PL::r0; n = PL::r0 * PL::r1; }
      else { const int e = i - N2; q = e / (PL::r0 * PL::r1) + 1; k = e % (PL::r0 * PL::r1); n = PL::r0 * PL::r1 * PL::r2; }
      float sn, cs;
      sincospif(2.0f * (float)(q * k) / (float)n, &sn, &cs);
      tw[i] = make_float2(cs, sn);
    }
  }
  // last (radix-2) stage: this thread owns k = 256 rank + kl for the rows of its parity
  const int kl = t & 255, rpar = t >> 8;
  const int kk = 256 * (int)rank + kl;
  float2 wk;
  sincospif((float)kk / 512.0f, &wk.y, &wk.x);                     // W_1024^k = e^{+j 2 pi k / 1024}
  __syncthreads();
  const int p = 2 * t + (int)rank;
  const long long cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const long long nspans = prm.spans_per_phase * prm.os;
  const long long rstride = (long long)prm.os * M;
  unsigned tile = 0;
  for (long long s = cid; s < nspans; s += ncl) {                  // cluster-uniform
    const Span sp = make_span(prm, s);
    if (sp.count <= 0) continue;
    const int rs = padi<MS>(((p - sp.shift + M) % M) >> 1);        // position in this CTA's half-sequence
    float2* gout = prm.out + (sp.m0 - prm.row_base) * (long long)M;
    fir_span<P, IN16, M, 0>(prm, sp, p, [&](int ii, long long i, float2 v) {
      buf0[(ii % RT) * S + rs] = v;
      if (ii % RT == RT - 1) {
        float2* rb = res + (size_t)(tile & 1) * RT * MS;
        __syncthreads();
        stockham_pass<MS, PL::r0, 1, RT, 512, false, false>(buf0, buf1, tw, nullptr, t, nullptr, 0, 0, 0);
        __syncthreads();
        stockham_pass<MS, PL::r1, PL::r0, RT, 512, false, false>(buf1, buf0, tw, nullptr, t, nullptr, 0, 0, 0);
        __syncthreads();
        stockham_pass<MS, PL::r2, PL::r0 * PL::r1, RT, 512, true, false>(buf0, buf1, tw, nullptr, t, rb, (long long)MS, 0, RT);
        cluster_barrier();                                         // both halves' results are in place
        const long long i0 = i - (RT - 1);
        const long long left = sp.count - i0;
        const int vhi = (int)(left < RT ? (left < 0 ? 0 : left) : RT);
        const int vlo = i0 < sp.skip ? (int)(sp.skip - i0) : 0;
        // all DSMEM loads of this thread's rows are issued before any is used (each costs ~200+ cycles)
        float2 ea[RT / 2], ob[RT / 2];
        #pragma unroll
        for (int u = 0; u < RT / 2; u++) {
          const float2* mine = rb + (rpar + 2 * u) * MS + kk;
          ea[u] = rank == 0 ? *mine : ld_dsmem(mine, 0);           // E[k]
          ob[u] = rank == 1 ? *mine : ld_dsmem(mine, 1);           // O[k]
        }
        #pragma unroll
        for (int u = 0; u < RT / 2; u++) {
          const int r = rpar + 2 * u;
          const float2 o = cmul(ob[u], wk);
          if (r >= vlo && r < vhi) {
            float2* g = gout + (i0 + r) * rstride + kk;
            g[0] = cadd(ea[u], o);
            g[MS] = csub(ea[u], o);
          }
        }
        tile++;
      }
    });
  }
  cluster_barrier();   // do not exit while the partner may still read this CTA's shared memory
}

// ---- large M, pipelined split path: ONE persistent launch, FIR tasks and FFT tasks from one queue -----
// The split path moves 4 + 8 + 8 + 8 B per sample through DRAM because a whole recording's FIR output
// is written before the row FFT reads it back.  Here both stages run inside one launch and the FFT
// trails the FIR by a few row groups, so the intermediate rows are still in the 126 MB L2 when they are
// transformed in place: DRAM sees the raw input once and the final rows once (the fused kernel's
// 4 + 8 B); the intermediate costs L2 bandwidth only.
//   * the recording is cut into row groups of os*span_rows rows.  FIR task = (group, phase, pair of
//     128-branch blocks): the register-window FIR of k_fir, two branch blocks side by side in a 256-thread
//     CTA.  FFT task = SUB consecutive rows of a group: the body of k_fft_rows_big, in place.
//   * tasks sit in ONE statically ordered queue: slot s = [FIR tasks of group s][FFT tasks of group s - lag].
//     A CTA draws tickets with one atomicAdd (the next ticket is requested while the current task runs).
//     An FFT task spins until the groups it reads are complete (per-group counters, release/acquire at gpu
//     scope); `lag` is chosen so that this wait is normally over before the ticket is drawn.
//   * deadlock-free without co-residency assumptions: a task only ever waits for tasks with smaller
//     tickets, and every CTA works through its tickets in increasing order.
// Arithmetic is exactly the split path's (same FIR pairs, same FFT plan): results are bit-identical.
struct PipeParams {
  unsigned long long* ticket;   // zeroed before the launch
  int* done;                    // [ngroups_fir] finished FIR tasks per group, zeroed before the launch
  int ngroups_fir;              // span groups (make_span's spans_per_phase)
  int ngroups_fft;              // row groups of os*span_rows rows that contain at least one row
  int lag;                      // slots between a group's FIR tasks and its FFT tasks
  int tpg;                      // FIR tasks per group: os * M / 256
  int tsub;                     // FFT tasks per group
  int sub_rows;                 // rows per FFT task (multiple of the FFT tile height)
  int need_next;                // a phase starts on an odd global row: group g's rows extend into span group g+1
  long long total;              // tickets
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int M, int P, bool IN16>
__global__ void __launch_bounds__(256, 2) k_chan_pipe(ChanParams prm, PipeParams pp) {
  typedef Plan<M> PL;
  static_assert(PL::np == 3 && PL::r0 == 16 && 4096 % M == 0, "large-M plan expected");
  constexpr int ROWS = 4096 / M, S = RowStride<M>::value, BPR0 = M / 16, NBB2 = M / 256;
  extern __shared__ float2 smem[];
  __shared__ long long s_ticket;
  float2* bufA = smem;
  float2* bufB = bufA + ROWS * S;
  const int t = threadIdx.x;
  float2* const y = prm.out;
  if (t == 0) s_ticket = (long long)atomicAdd(pp.ticket, 1ULL);
  __syncthreads();
  long long tk = s_ticket;
  const int slot_len = pp.tpg + pp.tsub;
  const long long group_rows = (long long)prm.os * prm.span_rows;
  while (tk < pp.total) {
    long long nxt = 0;
    if (t == 0) nxt = (long long)atomicAdd(pp.ticket, 1ULL);   // used only after the task: latency hidden
    const long long slot = tk / slot_len;
    const int idx = (int)(tk - slot * slot_len);
    if (idx < pp.tpg) {
      // ---------------- FIR task: span group `slot`, phase idx / NBB2, branch blocks 2*(idx % NBB2) + {0,1}
      if (slot < pp.ngroups_fir) {
        const int phase = idx / NBB2, pair = idx - phase * NBB2;
        const int p = (2 * pair + (t >> 7)) * 128 + (t & 127);
        const Span sp = make_span(prm, slot * prm.os + phase);
        if (sp.count > 0) {
          const int r = (p - sp.shift + M) % M;          // u'[r] = u[(r + shift) mod M]
          float2* dst = y + (sp.m0 - prm.row_base) * (long long)M + r;
          const long long rstride = (long long)prm.os * M;
          fir_span<P, IN16, M, 1>(prm, sp, p, [&](int, long long i, float2 v) {
            if (i >= sp.skip && i < sp.count) dst[i * rstride] = v;
          });
        }
        __syncthreads();                                 // every thread's rows are written ...
        if (t == 0) {
          __threadfence();                               // ... and ordered before the group's counter moves
          atomicAdd(pp.done + slot, 1);
        }
      }
    } else {
      // ---------------- FFT task: rows [g*group_rows + j*sub_rows, + sub_rows) of the call, in place
      const long long g = slot - pp.lag;
      if (g >= 0 && g < pp.ngroups_fft) {
        const int j = idx - pp.tpg;
        const long long gbeg = g * group_rows;
        long long r_begin = gbeg + (long long)j * pp.sub_rows;
        long long r_end = r_begin + pp.sub_rows;
        if (r_end > gbeg + group_rows) r_end = gbeg + group_rows;
        if (r_end > prm.nrows) r_end = prm.nrows;
        if (r_begin < r_end) {                           // block-uniform
          if (t == 0) {
            while (ld_acquire_gpu(pp.done + g) < pp.tpg) __nanosleep(64);
            if (pp.need_next && g + 1 < pp.ngroups_fir)
              while (ld_acquire_gpu(pp.done + g + 1) < pp.tpg) __nanosleep(64);
          }
          __syncthreads();
          const int row = t / BPR0, jj = t % BPR0;
          auto load = [&](long long r0, float2 (&v)[16]) {   // L2 loads: another SM wrote these rows
            const bool ok = r0 + row < r_end;
            const float2* src = y + (r0 + row) * (long long)M + jj;
            #pragma unroll
            for (int q = 0; q < 16; q++) v[q] = ok ? __ldcg(src + q * BPR0) : make_float2(0.f, 0.f);
          };
          float2 cur[16];
          load(r_begin, cur);
          for (long long r0 = r_begin; r0 < r_end; r0 += ROWS) {
            float2 nx[16];
            if (r0 + ROWS < r_end) load(r0 + ROWS, nx);
            const long long left = r_end - r0;
            const int vhi = (int)(left < ROWS ? left : ROWS);
            dft<16>(cur);
            {
              float2* d = bufA + row * S;
              #pragma unroll
              for (int q = 0; q < 16; q++) d[padi<M>(jj * 16 + q)] = cur[q];
            }
            __syncthreads();
            stockham_pass<M, PL::r1, PL::r0, ROWS, 256, false, false>(bufA, bufB, prm.tw, nullptr, t, nullptr, 0, 0, 0);
            __syncthreads();
            stockham_pass<M, PL::r2, PL::r0 * PL::r1, ROWS, 256, true, false>(bufB, bufA, prm.tw, nullptr, t,
                                                                              y + r0 * (long long)M, (long long)M, 0, vhi);
            #pragma unroll
            for (int q = 0; q < 16; q++) cur[q] = nx[q];
          }
        }
      }
    }
    __syncthreads();                                     // shared buffers and s_ticket are free again
    if (t == 0) s_ticket = nxt;
    __syncthreads();
    tk = s_ticket;
  }
}


}  // namespace chzi
