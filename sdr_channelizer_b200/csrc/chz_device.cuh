// Device-side building blocks shared by the channelizer kernels (sm_100a only).
//   K1  raw int8/int16 I/Q -> fp32 (exact; matlab/create_pdws_channelized.m:35-38)
//   K3  in-register DFT-2/4/8/16 and the shared-memory Stockham FFT over polyphase branches
#pragma once
#ifndef CHZ_TWREG
#define CHZ_TWREG 1
#endif
#include <cuda_runtime.h>
#include <stdint.h>

namespace chzi {

// ---- complex helpers (float2 = re,im).  __f*2_rn map to the packed FADD2/FMUL2/FFMA2 of sm_100 ----
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a*w in packed form: FMUL2 (a.x broadcast) + FFMA2 (a.y broadcast, w with its halves swapped and one sign
// flipped) = 3 FMA-pipe instructions including the sign flip; every FMA-pipe instruction, scalar or packed,
// occupies the pipe for two cycles on sm_100, so the scalar form below (2 FMUL + 2 FFMA) costs 4.
__device__ __forceinline__ float2 cmul_p(float2 a, float2 w) {
  return __ffma2_rn(make_float2(a.y, a.y), make_float2(-w.y, w.x), __fmul2_rn(make_float2(a.x, a.x), w));
}
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {   // a*w
#ifdef CHZ_CMUL_PACKED
  return cmul_p(a, w);
#else
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
#endif
}
__device__ __forceinline__ float2 mul_j(float2 a) { return make_float2(-a.y, a.x); }    // * (+j)

// ---- K1: int8/int16 pair -> two floats (exact; the 2^-(bitWidth-1) scale is applied by the caller
// or folded into the taps).  Two I2F conversions: they issue to the otherwise idle conversion (XU)
// pipe and keep the FMA pipe, which bounds this kernel, free.  Measured on B200 against a
// magic-number variant (XOR sign bit, PRMT into 0x4B00xxxx, packed FADD of -(2^23+2^15)), which cost
// one FADD2 per sample on the FMA pipe: 435.6 vs 421.0 GS/s on the fused M=64 kernel.
template <bool IN16> struct RawT;
template <> struct RawT<true> { typedef uint32_t type; };    // int16 I, int16 Q
template <> struct RawT<false> { typedef uint16_t type; };   // int8 I, int8 Q

template <bool IN16>
__device__ __forceinline__ float2 unpack_raw(uint32_t raw) {   // integer-valued, NOT yet scaled
  if (IN16) return make_float2((float)(short)(raw & 0xffffu), (float)(short)(raw >> 16));
  return make_float2((float)(signed char)(raw & 0xffu), (float)(signed char)((raw >> 8) & 0xffu));
}

// ---- in-register DFTs, exponent +j (y_t = sum_s v_s e^{+j 2 pi s t / R}), natural order in/out ----
__device__ __forceinline__ void dft2(float2& a, float2& b) {
  const float2 t = csub(a, b);
  a = cadd(a, b);
  b = t;
}
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_j(csub(a1, a3));
  a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}
__device__ __forceinline__ void dft8(float2* v) {
  // DIT: E = DFT4(v0,v2,v4,v6), O = DFT4(v1,v3,v5,v7); X[k] = E[k] + W8^k O[k], X[k+4] = E[k] - W8^k O[k]
  dft4(v[0], v[2], v[4], v[6]);
  dft4(v[1], v[3], v[5], v[7]);
  // W8^1 = (1+j)/sqrt2, W8^3 = (-1+j)/sqrt2: the 1/sqrt2 scale rides on the final add as an FMA
  const float c = 0.70710678118654752440f;
  const float2 cc = make_float2(c, c), nc = make_float2(-c, -c);
  const float2 t1 = make_float2(v[3].x - v[3].y, v[3].x + v[3].y);     // O[1] * (1+j)
  const float2 o2 = mul_j(v[5]);                                       // W8^2 = j
  const float2 t3 = make_float2(-(v[7].x + v[7].y), v[7].x - v[7].y);  // O[3] * (-1+j)
  const float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = __ffma2_rn(cc, t1, e1); v[5] = __ffma2_rn(nc, t1, e1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = __ffma2_rn(cc, t3, e3); v[7] = __ffma2_rn(nc, t3, e3);
}
__device__ __forceinline__ void dft16(float2* v) {
  // n = 4 n1 + n2, k = k1 + 4 k2:  A[n2][k1] = DFT4 over n1; twiddle W16^{n2 k1}; DFT4 over n2.
  #pragma unroll
  for (int n2 = 0; n2 < 4; n2++) dft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);   // v[n2 + 4 k1] = A[n2][k1]
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, c2 = 0.70710678118654752440f;
  // k1 = 1: W16^{n2}, n2 = 1,2,3
  v[5] = cmul(v[5], make_float2(c1, s1));
  v[6] = cmul(v[6], make_float2(c2, c2));
  v[7] = cmul(v[7], make_float2(s1, c1));
  // k1 = 2: W16^{2 n2} = W8^{n2}
  v[9] = cmul(v[9], make_float2(c2, c2));
  v[10] = mul_j(v[10]);
  v[11] = cmul(v[11], make_float2(-c2, c2));
  // k1 = 3: W16^{3 n2}
  v[13] = cmul(v[13], make_float2(s1, c1));
  v[14] = cmul(v[14], make_float2(-c2, c2));
  v[15] = cmul(v[15], make_float2(-c1, -s1));
  #pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);  // -> X[k1 + 4 k2] at v[4 k1 + k2]
  // reorder to natural k = k1 + 4 k2  (register renaming, no data movement after unrolling)
  float2 t[16];
  #pragma unroll
  for (int k1 = 0; k1 < 4; k1++)
    #pragma unroll
    for (int k2 = 0; k2 < 4; k2++) t[k1 + 4 * k2] = v[4 * k1 + k2];
  #pragma unroll
  for (int i = 0; i < 16; i++) v[i] = t[i];
}
// Odd prime radices for the reference's natural channel counts (M = fs*1e-6 = 56 = 8*7,
// matlab/create_pdws_channelized.m:31; 560 = 16*5*7, generate_channelized_training_iq.m:95-96):
//   X_k = v_0 + sum_{m=1..(R-1)/2} [ (v_m + v_{R-m}) cos(2 pi k m / R) + j (v_m - v_{R-m}) sin(2 pi k m / R) ]
// and X_{R-k} is the same with the second term negated.  Indices are compile-time after unrolling, so
// the cosines/sines fold into immediates.
__host__ __device__ constexpr float cos_r(int R, int i) {   // cos(2 pi i / R), R in {3, 5, 7}, 0 <= i < R
  return R == 3 ? (i == 0 ? 1.0f : -0.5f) : R == 5 ? (i == 0 ? 1.0f : (i == 1 || i == 4) ? 0.30901699437494742410f : -0.80901699437494742410f)
                : (i == 0 ? 1.0f : (i == 1 || i == 6) ? 0.62348980185873353053f
                   : (i == 2 || i == 5) ? -0.22252093395631440429f : -0.90096886790241912624f);
}
__host__ __device__ constexpr float sin_r(int R, int i) {   // sin(2 pi i / R)
  return R == 3 ? (i == 0 ? 0.0f : i == 1 ? 0.86602540378443864676f : -0.86602540378443864676f) : R == 5 ? (i == 0 ? 0.0f : i == 1 ? 0.95105651629515357212f : i == 2 ? 0.58778525229247312917f
                   : i == 3 ? -0.58778525229247312917f : -0.95105651629515357212f)
                : (i == 0 ? 0.0f : i == 1 ? 0.78183148246802980871f : i == 2 ? 0.97492791218182360702f
                   : i == 3 ? 0.43388373911755812048f : i == 4 ? -0.43388373911755812048f
                   : i == 5 ? -0.97492791218182360702f : -0.78183148246802980871f);
}
template <int R> __device__ __forceinline__ void dft_odd(float2* v) {
  constexpr int H = (R - 1) / 2;
  float2 a[H], b[H];
  #pragma unroll
  for (int m = 1; m <= H; m++) { a[m - 1] = cadd(v[m], v[R - m]); b[m - 1] = mul_j(csub(v[m], v[R - m])); }
  float2 x0 = v[0];
  #pragma unroll
  for (int m = 0; m < H; m++) x0 = cadd(x0, a[m]);
  float2 out[R];
  out[0] = x0;
  #pragma unroll
  for (int k = 1; k <= H; k++) {
    float2 re = v[0], im = make_float2(0.f, 0.f);
    #pragma unroll
    for (int m = 1; m <= H; m++) {
      const float c = cos_r(R, (k * m) % R), sn = sin_r(R, (k * m) % R);
      re = __ffma2_rn(make_float2(c, c), a[m - 1], re);
      im = __ffma2_rn(make_float2(sn, sn), b[m - 1], im);
    }
    out[k] = cadd(re, im);
    out[R - k] = csub(re, im);
  }
  #pragma unroll
  for (int i = 0; i < R; i++) v[i] = out[i];
}
template <int R> __device__ __forceinline__ void dft(float2* v);
template <> __device__ __forceinline__ void dft<3>(float2* v) { dft_odd<3>(v); }
template <> __device__ __forceinline__ void dft<5>(float2* v) { dft_odd<5>(v); }
template <> __device__ __forceinline__ void dft<7>(float2* v) { dft_odd<7>(v); }
template <> __device__ __forceinline__ void dft<2>(float2* v) { dft2(v[0], v[1]); }
template <> __device__ __forceinline__ void dft<4>(float2* v) { dft4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ void dft<8>(float2* v) { dft8(v); }
template <> __device__ __forceinline__ void dft<16>(float2* v) { dft16(v); }

// ---- FFT plan: radices per pass.  First pass radix is 16 or 8 so that every later Stockham pass
// writes runs of >= 16 contiguous elements (bank-conflict-free), see DESIGN.md. -----------------------
template <int M> struct Plan;
template <> struct Plan<8>    { static constexpr int np = 1; static constexpr int r0 = 8,  r1 = 1,  r2 = 1; };
template <> struct Plan<16>   { static constexpr int np = 1; static constexpr int r0 = 16, r1 = 1,  r2 = 1; };
template <> struct Plan<32>   { static constexpr int np = 2; static constexpr int r0 = 8,  r1 = 4,  r2 = 1; };
template <> struct Plan<64>   { static constexpr int np = 2; static constexpr int r0 = 8,  r1 = 8,  r2 = 1; };
template <> struct Plan<128>  { static constexpr int np = 2; static constexpr int r0 = 16, r1 = 8,  r2 = 1; };
template <> struct Plan<256>  { static constexpr int np = 2; static constexpr int r0 = 16, r1 = 16, r2 = 1; };
template <> struct Plan<512>  { static constexpr int np = 3; static constexpr int r0 = 16, r1 = 8,  r2 = 4; };
template <> struct Plan<1024> { static constexpr int np = 3; static constexpr int r0 = 16, r1 = 8,  r2 = 8; };
template <> struct Plan<2048> { static constexpr int np = 3; static constexpr int r0 = 16, r1 = 16, r2 = 8; };
template <> struct Plan<4096> { static constexpr int np = 3; static constexpr int r0 = 16, r1 = 16, r2 = 16; };
template <> struct Plan<56>   { static constexpr int np = 2; static constexpr int r0 = 8,  r1 = 7,  r2 = 1; };   // fs = 56 MS/s, 1 MHz bins
template <> struct Plan<560>  { static constexpr int np = 3; static constexpr int r0 = 16, r1 = 5,  r2 = 7; };   // 0.1 MHz bins

// Threads of one group in the fused kernel: a branch per thread, rounded up to whole warps for M >= 32
// (named barriers count whole warps); the spare lanes of a non-power-of-two M idle through the FIR and
// work in the FFT passes.
template <int M> struct GroupThreads { static constexpr int value = M < 32 ? M : (M + 31) / 32 * 32; };

// Padded element index inside one row of the shared tile (8-byte elements, 16 per 128-byte bank
// row).  One pad element per 2^PadShift elements, and a row stride chosen so that the half-warps of
// every pass (a few butterflies of several consecutive rows) hit 16 distinct 8-byte bank pairs both
// when they read with stride M/R and when they write the Stockham-transposed result:
//   M = 32 (radix 8,4):  4 butterflies per row  -> pad every 8, rows 4 bank pairs apart (stride 36)
//   M = 64 (radix 8,8):  8 butterflies per row  -> pad every 8, rows 8 apart           (stride 72)
//   M >= 128 (first radix 16): pad every 16, rows 8 apart for M = 128, irrelevant above
template <int M> struct PadShift { static constexpr int value = (M == 32 || M == 64 || M == 56) ? 3 : 4; };
template <int M> __device__ __forceinline__ int padi(int i) { return i + (i >> PadShift<M>::value); }
// Layout of the tile the FIRST pass reads (the one the FIR threads, or a staging copy, write with one lane per
// branch).  With a pad every 8 elements the 16 lanes of a half-warp store to slots {0..7, 9..16}: slot 16 falls on
// the banks of slot 0, so every FIR store costs two wavefronts per half-warp instead of one (ncu on M = 64: 38.4 M of
// 192 M shared wavefronts).  The first pass does not need that pad: it reads BPR consecutive elements per row and
// RowStride puts consecutive rows half (M = 64, 56) or a quarter (M = 32) of a 128-byte bank window apart, so plain
// rows are conflict free for the stores and for the reads.  Measured on B200: M = 32 84.0 -> 85.2 %, M = 56 73.7 ->
// 74.4 % of the HBM roofline; at M = 64 the conflicts drop from 23 % to 6 % of the wavefronts and the LSU data pipe
// from 82 % to 70 % busy, but the kernel gets 3.5 % SLOWER at 16 taps per band (13 M more instructions after
// re-scheduling around its 128-register limit; unchanged at 12 taps) -- the LSU pipe was not what bounds it
// (profiles/r02g_*).  So the plain layout is used for 32 and 56 only.
template <int M> struct FirstPassPlain { static constexpr bool value = (M == 32 || M == 56); };
template <int M> __device__ __forceinline__ int padi_first(int i) { return FirstPassPlain<M>::value ? i : padi<M>(i); }
template <int M> struct RowStride {
  static constexpr int value = M == 32 ? 36 : (M == 64 ? 72 : (M == 56 ? 72 : M + M / 16 + (M < 16 ? 1 : 0)));
};

// One Stockham pass of radix R over ROWS rows of length M held in shared memory, NT threads.
//   src/dst : [ROWS][RowStride<M>] float2 (padded with padi)
//   NS      : product of the radices of earlier passes
//   LAST    : write to global memory (gout + row*grow_stride + k) instead of dst; only rows in [vlo, vhi)
//   TWREG   : inter-pass twiddles come from twr[q-1] (registers, loaded once per thread by the caller;
//             valid when NT is a multiple of M/R so a thread always owns the same butterfly column)
// Everything about the geometry is a compile-time constant, so after unrolling each access is one
// per-thread base address plus an immediate offset.
template <int M, int R, int NS, int ROWS, int NT, bool LAST, bool TWREG>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ src, float2* __restrict__ dst,
                                              const float2* __restrict__ tw, const float2* twr, int t,
                                              float2* __restrict__ gout, long long grow_stride, int vlo, int vhi) {
  constexpr int S = RowStride<M>::value;
  constexpr int BPR = M / R;            // butterflies per row
  constexpr int TOTAL = ROWS * BPR;
  constexpr int ITERS = (TOTAL + NT - 1) / NT;
  #pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int b = t + it * NT;
    if (TOTAL % NT != 0 && b >= TOTAL) break;
    const int row = b / BPR, j = b % BPR;
    const float2* s = src + row * S;
    float2 v[R];
    #pragma unroll
    for (int q = 0; q < R; q++) v[q] = s[NS == 1 ? padi_first<M>(j + q * BPR) : padi<M>(j + q * BPR)];
    const int k = j % NS;               // NS is a compile-time constant (a power of two except in Plan<560>)
    if (NS > 1) {
      // twiddle table layout (host: build_twiddles): per pass, entry (q-1)*NS + k holds
      // W_{NS R}^{q k}; the lanes of a warp read consecutive k -> consecutive addresses, no conflicts
      constexpr int TWOFF = (NS == Plan<M>::r0) ? 0 : (Plan<M>::r1 - 1) * Plan<M>::r0;
      #pragma unroll
      for (int q = 1; q < R; q++) v[q] = cmul(v[q], TWREG ? twr[q - 1] : tw[TWOFF + (q - 1) * NS + k]);
    }
    dft<R>(v);
    const int j0 = (j - k) * R + k;     // (j / NS) * NS * R + k
    if (LAST) {
      // Tried: swapping every other result with the neighbouring lane (adjacent lanes hold adjacent channels) so that
      // each lane stores 16 bytes -- 512-byte runs per warp instruction.  A do-nothing kernel with this traffic mix
      // runs 7 % faster with 8-byte loads / 16-byte stores than with 4 / 8 (tools/ubench/mixbw.cu), but here the two
      // shuffles per pair land on the LSU pipe, the busiest unit of the fused kernels: M = 64 442.8 -> 360.6 GS/s,
      // M = 32 459.7 -> 324.7, M = 256 352.5 -> 327.5 (profiles/r02o_paired_stores_ab.jsonl).  Removed.
      if (row >= vlo && row < vhi) {
        float2* g = gout + (long long)row * grow_stride;
        #pragma unroll
        for (int q = 0; q < R; q++) g[j0 + q * NS] = v[q];
      }
    } else {
      float2* d = dst + row * S;
      #pragma unroll
      for (int q = 0; q < R; q++) d[padi<M>(j0 + q * NS)] = v[q];
    }
  }
}

// Can the last pass keep its twiddles in registers?  (two-pass plans with a small last radix)
template <int M, int NT> struct TwReg {
  typedef Plan<M> PL;
  static constexpr int RL = PL::np == 2 ? PL::r1 : 1;                 // radix of the last pass
  static constexpr bool value = CHZ_TWREG && PL::np == 2 && RL <= 8 && (NT % (M / RL) == 0);
  static constexpr int count = value ? RL - 1 : 1;
};
// Per-thread twiddles of the last pass of a two-pass plan: W_{r0 r1}^{q (j mod r0)}, q = 1..r1-1
template <int M, int NT>
__device__ __forceinline__ void load_last_pass_twiddles(const float2* __restrict__ tw_g, int t, float2* twr) {
  typedef Plan<M> PL;
  if constexpr (TwReg<M, NT>::value) {
    constexpr int R = PL::r1, NS = PL::r0, BPR = M / R;
    const int k = (t % BPR) & (NS - 1);
    #pragma unroll
    for (int q = 1; q < R; q++) twr[q - 1] = tw_g[(q - 1) * NS + k];
  }
}

// Full M-point FFT (exponent +j) of ROWS rows by NT threads.  buf0 holds the input (padded layout);
// buf1 is scratch of the same size.  Result goes to global memory in natural order.  `sync()` must
// synchronise exactly the threads that cooperate on this tile.
template <int M, int ROWS, int NT, bool TWREG, typename SyncF>
__device__ __forceinline__ void fft_tile_to_global(float2* buf0, float2* buf1, const float2* tw, const float2* twr,
                                                   int t, float2* gout, long long grow_stride, int vlo, int vhi,
                                                   SyncF sync) {
  typedef Plan<M> PL;
  if constexpr (PL::np == 1) {
    // M = 8, 16: one butterfly IS a whole row, so a thread that stored its own results would write 8-byte
    // pieces of 32 different rows per instruction (partial sectors: 4x the L1->L2 write traffic; measured
    // 40 % / 29 % of the HBM roofline).  The rows go through buf1 instead and are written out with the
    // lanes across channels: every instruction stores whole 64- or 128-byte rows.
    stockham_pass<M, PL::r0, 1, ROWS, NT, false, false>(buf0, buf1, tw, twr, t, gout, grow_stride, vlo, vhi);
    sync();
    constexpr int S = RowStride<M>::value;
    if constexpr (NT == M) {   // fused kernel: thread t owns channel t of every row of the tile
      #pragma unroll
      for (int row = 0; row < ROWS; row++)
        if (row >= vlo && row < vhi) gout[(long long)row * grow_stride + t] = buf1[row * S + padi<M>(t)];
    } else {
      #pragma unroll
      for (int e0 = 0; e0 < ROWS * M; e0 += NT) {
        const int e = e0 + t;
        if ((ROWS * M) % NT != 0 && e >= ROWS * M) break;
        const int row = e / M, i = e % M;
        if (row >= vlo && row < vhi) gout[(long long)row * grow_stride + i] = buf1[row * S + padi<M>(i)];
      }
    }
  } else if constexpr (PL::np == 2) {
    stockham_pass<M, PL::r0, 1, ROWS, NT, false, false>(buf0, buf1, tw, twr, t, gout, grow_stride, vlo, vhi);
    sync();
    stockham_pass<M, PL::r1, PL::r0, ROWS, NT, true, TWREG>(buf1, buf0, tw, twr, t, gout, grow_stride, vlo, vhi);
  } else {
    stockham_pass<M, PL::r0, 1, ROWS, NT, false, false>(buf0, buf1, tw, twr, t, gout, grow_stride, vlo, vhi);
    sync();
    stockham_pass<M, PL::r1, PL::r0, ROWS, NT, false, false>(buf1, buf0, tw, twr, t, gout, grow_stride, vlo, vhi);
    sync();
    stockham_pass<M, PL::r2, PL::r0 * PL::r1, ROWS, NT, true, false>(buf0, buf1, tw, twr, t, gout, grow_stride, vlo, vhi);
  }
}

}  // namespace chzi
