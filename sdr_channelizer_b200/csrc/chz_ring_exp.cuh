// Ring kernels that lost their A/B against k_chan_ring_ws (chz_ring.cuh); built only with `make EXPERIMENTS=1`
// and selected with CHZ_RING_VARIANT (1 = every warp does FIR then FFT, 2 = 1024 threads with one branch each).
#pragma once
#include "chz_ring.cuh"

namespace chzi {
namespace ring {

// ---- single-role variant (CHZ_RING_VARIANT=1): the first round-2 kernel.  512 threads, two adjacent branches each;
// every warp filters its columns of a tile, then the CTA transforms it.  273 / 136 GS/s (critical / 2x, P = 16)
// against 310 / 162 for the role-split kernel: profiles/r02_ring_ncu.txt, r02j_ring_ws_ab.jsonl.
template <int P, bool IN16, int UNPACK>
__global__ void __launch_bounds__(kNT, 1) k_chan_ring(ChanParams prm, RingParams rp) {
  typedef Smem<IN16> SM;
  typedef typename RawT<IN16>::type raw_t;
  constexpr int M = kM, ROWB = SM::ROWB, SLOTB = SM::SLOTB, TS = kTileStride;
  constexpr int J0 = 16 - P;                      // first ring row (relative to a0-16) a delta = 0 thread reads
  extern __shared__ __align__(128) unsigned char smem[];
  float2* tiles = (float2*)(smem + SM::OFF_TILE);
  float* h0s = (float*)(smem + SM::OFF_H0);
  float2* tw1s = (float2*)(smem + SM::OFF_TW1);
  const unsigned ring_s = smem_u32(smem), bar = smem_u32(smem + SM::OFF_BAR);
  const int t = threadIdx.x;
  const int os = prm.os, D = prm.D;

  // ---- this CTA's run of steps ----
  const long long k0 = rp.nsteps * blockIdx.x / gridDim.x, k1 = rp.nsteps * (blockIdx.x + 1) / gridDim.x;
  if (k0 >= k1) return;

  // ---- persistent per-thread state: taps of branches 2t+1 and (2t+2) mod M, twiddles of passes 0 and 1 ----
  const int b1 = 2 * t + 1, b2 = (2 * t + 2) & (M - 1);
  float h1[P], h2[P];
  #pragma unroll
  for (int q = 0; q < P; q++) { h1[q] = __ldg(prm.taps + q * M + b1); h2[q] = __ldg(prm.taps + q * M + b2); }
  if (t < P) h0s[t] = __ldg(prm.taps + t * M);
  float2 tw0[7];
  #pragma unroll
  for (int k = 1; k < 8; k++) tw0[k - 1] = __ldg(rp.twn + (((t & 127) * k) & (M - 1)));   // W_M^{j k}, j = t mod 128
  if (t < 7 * 16) tw1s[t] = __ldg(rp.twn + ((((t & 15) * ((t >> 4) + 1)) << 3) & (M - 1)));   // W_128^{j1 k} at [k-1][j1]
  if (t == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long in_end = prm.in_base + prm.n_in;
  // bulk copies need 16-byte aligned global addresses: frame starts are multiples of M samples from in_base
  const bool aligned = ((((unsigned long long)prm.in) - (unsigned long long)prm.in_base * sizeof(raw_t)) & 15ull) == 0;
  const raw_t* __restrict__ inp = (const raw_t*)prm.in - prm.in_base;   // inp[idx], idx in [in_base, in_end)
  unsigned parity = 0;
  bool pending = false;                           // a bulk copy into the newest slot is in flight

  // frames [a, a+8) -> ring slot s.  Whole slot inside this call's input and aligned: one bulk copy issued by
  // thread 0 (the caller waits on the mbarrier before reading); otherwise every thread copies with bounds
  // checks (history buffer, zeros before the stream start and past the data) and the caller synchronises.
  auto load_slot = [&](long long a, int s) -> bool {
    const long long lo = a * M, hi = lo + (long long)kR * M;
    if (rp.dbg & 4) return false;
    if (aligned && lo >= prm.in_base && hi <= in_end) {
      if (t == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, SLOTB);
        bulk_g2s(ring_s + s * SLOTB, inp + lo, SLOTB, bar);
      }
      return true;
    }
    raw_t* dst = (raw_t*)(smem + s * SLOTB);
    #pragma unroll 4
    for (int e = t; e < kR * M; e += kNT) dst[e] = (raw_t)load_raw<IN16>(prm, lo + e);
    return false;
  };

  const long long a_start = rp.a_lo + k0 * kR;
  // warm-up: the 16 frames before the first step.  One bulk copy at a time on the single mbarrier, and a
  // CTA barrier after every wait so that no thread can still be polling phase n when phase n + 1 completes.
  if (load_slot(a_start - 2 * kR, 0)) { mbar_wait(bar, parity); parity ^= 1; }
  __syncthreads();
  if (load_slot(a_start - kR, 1)) { mbar_wait(bar, parity); parity ^= 1; }
  __syncthreads();
  pending = load_slot(a_start, 2);

  __syncthreads();                                 // ring writes of a slow-path warm-up are visible

  // Per phase: FIR -> tile[buf] | CTA barrier | pass 0 | pass 1 | pass 2 + global stores -> straight into the next
  // phase's FIR, which writes the OTHER tile buffer.  Rows are independent in the FFT, so its passes are separated by
  // named barriers of the four warps that own a pair of rows only; one CTA-wide barrier remains per eight rows, and a
  // warp that is done with its rows starts filtering while others still stream theirs out.
  int s_old = 0;                                   // slot of frames a0-16 .. a0-9
  int buf = 0;
  for (long long k = k0; k < k1; k++) {
    const long long a0 = rp.a_lo + k * kR;
    const int s_mid = s_old == 2 ? 0 : s_old + 1, s_new = s_mid == 2 ? 0 : s_mid + 1;
    if (pending) { mbar_wait(bar, parity); parity ^= 1; }
    else __syncthreads();       // the newest slot was filled by every thread's bounds-checked copy: after the FIR's barrier only
                                // row-group barriers follow, so a CTA-wide one is needed before those writes are read
    const unsigned sb0 = ring_s + s_old * SLOTB, sb1 = ring_s + s_mid * SLOTB, sb2 = ring_s + s_new * SLOTB;

    for (int ph = 0; ph < os; ph++, buf ^= 1) {
      float2* tile = tiles + buf * (kR * TS);
      // ---- FIR: 8 rows x 2 branches per thread ----
      // Row m = os*a + ph, branch p reads x[a M + ph D - q M - p] = frame (a - q - 1 + delta), column cl:
      //   ph = 0: cl = M - p, delta = 0 (p >= 1);  ph = 1: p <= D: cl = D - p, delta = 1;  p > D: cl = M + D - p, delta = 0.
      // The pair (2t+2, 2t+1) is the 8-byte aligned pair of columns (cl, cl + 1).  Branch 0 (thread 511's
      // first element) is the one column whose delta differs from its neighbour's: fixed up below.
      const bool lowhalf = ph && t < D / 2;
      const int cl = ph ? (lowhalf ? D - 2 * t - 2 : M + D - 2 * t - 2) : M - 2 * t - 2;
      const unsigned dcol = (unsigned)cl * sizeof(raw_t) + (lowhalf ? ROWB : 0);
      const unsigned r0b = sb0 + dcol, r1b = sb1 + dcol, r2b = sb2 + dcol;
      const unsigned e0 = lowhalf ? sb1 + cl * (unsigned)sizeof(raw_t) : sb0 + 7 * ROWB + cl * (unsigned)sizeof(raw_t);
      const unsigned e1 = lowhalf ? sb2 + cl * (unsigned)sizeof(raw_t) : sb1 + 7 * ROWB + cl * (unsigned)sizeof(raw_t);
      if (!(rp.dbg & 2)) {
        float2 acc1[kR], acc2[kR];
        #pragma unroll
        for (int r = 0; r < kR; r++) { acc1[r] = make_float2(0.f, 0.f); acc2[r] = make_float2(0.f, 0.f); }
        #pragma unroll
        for (int ii = 0; ii < P + kR - 1; ii++) {
          const int j = ii + J0;                       // ring row (before delta), compile time
          const unsigned addr = (j & 7) == 7 ? (j < 8 ? e0 : e1) : ((j < 8 ? r0b : (j < 16 ? r1b : r2b)) + (j & 7) * ROWB);
          uint32_t wa, wb;                             // columns cl (branch 2t+2) and cl+1 (branch 2t+1)
          if (IN16) {
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wa), "=r"(wb) : "r"(addr));
          } else {
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr));
            wa = w & 0xffffu; wb = w >> 16;
          }
          const float2 xa = unpack<IN16, UNPACK>(wa), xb = unpack<IN16, UNPACK>(wb);
          #pragma unroll
          for (int r = 0; r < kR; r++) {
            const int q = r + P - 1 - ii;
            if (q >= 0 && q < P) acc2[r] = __ffma2_rn(make_float2(h2[q], h2[q]), xa, acc2[r]);
          }
          #pragma unroll
          for (int r = 0; r < kR; r++) {
            const int q = r + P - 1 - ii;
            if (q >= 0 && q < P) acc1[r] = __ffma2_rn(make_float2(h1[q], h1[q]), xb, acc1[r]);   // (one branch packed, one scalar: 243 against 273 GS/s)
          }
        }
        const int shift = ph ? D : 0;
        const int pos1 = tpad((b1 - shift) & (M - 1)), pos2 = tpad((b2 - shift) & (M - 1));
        #pragma unroll
        for (int r = 0; r < kR; r++) { tile[r * TS + pos1] = acc1[r]; tile[r * TS + pos2] = acc2[r]; }
        if (t >= kNT - 32) {
          // branch 0: u_0[m] = sum_q h[qM] x[a M + ph D - q M] = frame (a - q), column ph*D: rows 16 + r - q of the ring
          __syncwarp();
          const int r = t & 31;
          if (r < kR) {
            float2 acc = make_float2(0.f, 0.f);
            const unsigned c0 = (unsigned)(ph ? D : 0) * sizeof(raw_t);
            #pragma unroll
            for (int q = P - 1; q >= 0; q--) {
              const int i = 16 + r - q;                 // 1 .. 23
              const unsigned sb = i < 8 ? sb0 : (i < 16 ? sb1 : sb2);
              uint32_t w;
              if (IN16) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(sb + (i & 7) * ROWB + c0));
              else asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w) : "r"(sb + (i & 7) * ROWB + c0));
              acc = __ffma2_rn(make_float2(h0s[q], h0s[q]), unpack<IN16, UNPACK>(w), acc);
            }
            tile[r * TS + tpad((0 - shift) & (M - 1))] = acc;
          }
        }
      }
      __syncthreads();                               // barrier A: the tile is complete, nobody reads the ring any more
      // the oldest slot is dead after the last phase's FIR: request the next step's frames into it now, the
      // copy lands while the FFT passes run
      if (ph == os - 1) pending = (k + 1 < k1) ? load_slot(a0 + kR, s_old) : false;

      if (rp.dbg & 1) continue;
      // ---- FFT, decimation in frequency, in place: 8 (stride 128) x 8 (stride 16) x 16 (contiguous) ----
      {   // pass 0: z_{k0}[j] = W_M^{j k0} sum_q u[j + 128 q] W_8^{q k0}  ->  position 128 k0 + j
        const int j = t & 127, rr = t >> 7;
        // both butterflies of the thread are loaded before either is computed: twice the loads in flight per warp
        // (4 warps per scheduler is all the latency hiding this kernel has)
        float2* row0 = tile + rr * TS + j;
        float2* row1 = row0 + 4 * TS;
        float2 v[8], w[8];
        #pragma unroll
        for (int q = 0; q < 8; q++) { v[q] = row0[q * 130]; w[q] = row1[q * 130]; }
        dft8(v);
        dft8(w);
        #pragma unroll
        for (int q = 1; q < 8; q++) { v[q] = cmul(v[q], tw0[q - 1]); w[q] = cmul(w[q], tw0[q - 1]); }
        #pragma unroll
        for (int q = 0; q < 8; q++) { row0[q * 130] = v[q]; row1[q * 130] = w[q]; }
      }
      // the rest of the FFT is local to a pair of rows: rows rr and rr + 4 belong to the four warps t >> 7
      asm volatile("bar.sync %0, 128;" ::"r"((t >> 7) + 1) : "memory");
      {   // pass 1 inside block k0: w_{k1}[j1] = W_128^{j1 k1} sum_q z[j1 + 16 q] W_8^{q k1}  ->  position 128 k0 + 16 k1 + j1
        const int j1 = t & 15, kb = (t >> 4) & 7, rr = t >> 7;
        float2* row0 = tile + rr * TS + kb * 130 + j1;
        float2* row1 = row0 + 4 * TS;
        float2 v[8], w[8], tw[7];
        #pragma unroll
        for (int q = 0; q < 8; q++) { v[q] = row0[q * 16]; w[q] = row1[q * 16]; }
        #pragma unroll
        for (int q = 1; q < 8; q++) tw[q - 1] = tw1s[(q - 1) * 16 + j1];
        dft8(v);
        dft8(w);
        #pragma unroll
        for (int q = 1; q < 8; q++) { v[q] = cmul(v[q], tw[q - 1]); w[q] = cmul(w[q], tw[q - 1]); }
        #pragma unroll
        for (int q = 0; q < 8; q++) { row0[q * 16] = v[q]; row1[q * 16] = w[q]; }
      }
      asm volatile("bar.sync %0, 128;" ::"r"((t >> 7) + 1) : "memory");
      const int row_i = (t >> 7) + 4 * ((t >> 6) & 1), b = t & 63;
      {   // pass 2: y[k0 + 8 k1 + 64 k2] = sum_{j1} w[j1] W_16^{j1 k2}; lanes run over (k0, k1): 32 consecutive channels per store
        const int kb = b & 7, kc = b >> 3;
        const float4* src = (const float4*)(tile + row_i * TS + kb * 130 + kc * 16);
        float2 v[16];
        #pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 f = src[q];
          v[2 * q] = make_float2(f.x, f.y); v[2 * q + 1] = make_float2(f.z, f.w);
        }
        dft16(v);
        const long long m = (a0 + row_i) * os + ph;
        if (m >= prm.row_base && m < prm.row_base + prm.nrows) {
          float2* g = prm.out + (m - prm.row_base) * (long long)M + b;
          #pragma unroll
          for (int q = 0; q < 16; q++) g[q * 64] = v[q];
        }
      }
    }
    s_old = s_mid;
  }
}

// ---- 1024-thread variant: one branch per thread (make EXPERIMENTS=1, CHZ_RING_VARIANT=2) --------------------------
// MEASURED SLOWER than the 512-thread kernel: 225 against 273 GS/s critically sampled, 115.7 against 136.0 GS/s on
// configs[2] (profiles/r02h_ring_1024_threads_ab.jsonl).  Twice the warps do not buy latency hiding here: per output
// the addressing, the ring reads (LDS.32 per branch instead of LDS.64 per pair) and the re-read taps cost more issue
// slots than the shorter stalls give back.  Kept as a record of the experiment.
// Same ring, same tiles, same arithmetic per row (identical results), but 32 warps instead of 16: the 512-thread kernel
// alternates between an FMA-bound FIR and shared-memory-bound FFT passes with 4 warps per scheduler, i.e. with little
// latency hiding inside either.  64 registers per thread suffice because nothing stays resident between the phases: a
// thread owns ONE ring column (taps of its branch are re-read from L2 for every phase: 64 KB per CTA and phase, the
// pass-0 twiddles likewise), every radix-8 pass has exactly one butterfly per thread, and the last (radix-16) pass
// occupies the lower half of every row's warps while the upper half already filters the next phase into the other
// tile buffer.  A thread's single column also removes the branch-0 fix-up: its frame offset delta is per thread.
constexpr int kNT1k = 1024;

template <int P, bool IN16, int UNPACK>
__global__ void __launch_bounds__(kNT1k, 1) k_chan_ring1k(ChanParams prm, RingParams rp) {
  typedef Smem<IN16> SM;
  typedef typename RawT<IN16>::type raw_t;
  constexpr int M = kM, ROWB = SM::ROWB, SLOTB = SM::SLOTB, TS = kTileStride;
  constexpr int J0 = 16 - P;
  extern __shared__ __align__(128) unsigned char smem[];
  float2* tiles = (float2*)(smem + SM::OFF_TILE);
  float2* tw1s = (float2*)(smem + SM::OFF_TW1);
  const unsigned ring_s = smem_u32(smem), bar = smem_u32(smem + SM::OFF_BAR);
  const int t = threadIdx.x;
  const int os = prm.os, D = prm.D;
  const long long k0 = rp.nsteps * blockIdx.x / gridDim.x, k1 = rp.nsteps * (blockIdx.x + 1) / gridDim.x;
  if (k0 >= k1) return;

  const int c = t;                                  // ring column of phase 0; branch p0 = (M - c) mod M
  const int p0 = (M - c) & (M - 1);
  if (t < 7 * 16) tw1s[t] = __ldg(rp.twn + ((((t & 15) * ((t >> 4) + 1)) << 3) & (M - 1)));   // W_128^{j1 k} at [k-1][j1]
  if (t == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long in_end = prm.in_base + prm.n_in;
  const bool aligned = ((((unsigned long long)prm.in) - (unsigned long long)prm.in_base * sizeof(raw_t)) & 15ull) == 0;
  const raw_t* __restrict__ inp = (const raw_t*)prm.in - prm.in_base;
  unsigned parity = 0;
  bool pending = false;
  auto load_slot = [&](long long a, int s) -> bool {
    const long long lo = a * M, hi = lo + (long long)kR * M;
    if (aligned && lo >= prm.in_base && hi <= in_end) {
      if (t == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, SLOTB);
        bulk_g2s(ring_s + s * SLOTB, inp + lo, SLOTB, bar);
      }
      return true;
    }
    raw_t* dst = (raw_t*)(smem + s * SLOTB);
    #pragma unroll 4
    for (int e = t; e < kR * M; e += kNT1k) dst[e] = (raw_t)load_raw<IN16>(prm, lo + e);
    return false;
  };
  const long long a_start = rp.a_lo + k0 * kR;
  if (load_slot(a_start - 2 * kR, 0)) { mbar_wait(bar, parity); parity ^= 1; }
  __syncthreads();
  if (load_slot(a_start - kR, 1)) { mbar_wait(bar, parity); parity ^= 1; }
  __syncthreads();
  pending = load_slot(a_start, 2);
  __syncthreads();

  int s_old = 0, buf = 0;
  for (long long k = k0; k < k1; k++) {
    const long long a0 = rp.a_lo + k * kR;
    const int s_mid = s_old == 2 ? 0 : s_old + 1, s_new = s_mid == 2 ? 0 : s_mid + 1;
    if (pending) { mbar_wait(bar, parity); parity ^= 1; }
    else __syncthreads();       // the newest slot was filled by every thread's bounds-checked copy: after the FIR's barrier only
                                // row-group barriers follow, so a CTA-wide one is needed before those writes are read
    const unsigned sb0 = ring_s + s_old * SLOTB, sb1 = ring_s + s_mid * SLOTB, sb2 = ring_s + s_new * SLOTB;

    for (int ph = 0; ph < os; ph++, buf ^= 1) {
      float2* tile = tiles + buf * (kR * TS);
      {
        // ---- FIR: 8 rows of branch p0.  Row m = os*a + ph reads x[a M + ph D - q M - p0] = frame (a - q - 1 + delta),
        // column cc:  ph = 0: cc = c, delta = (c == 0);  ph = 1: cc = (c + D) mod M, delta = (p0 <= D)
        const int cc = ph ? ((c + D) & (M - 1)) : c;
        const bool dl = ph ? (p0 <= D) : (c == 0);
        const unsigned cb = (unsigned)cc * sizeof(raw_t);
        const unsigned dcol = cb + (dl ? ROWB : 0);
        const unsigned r0b = sb0 + dcol, r1b = sb1 + dcol, r2b = sb2 + dcol;
        const unsigned e0 = dl ? sb1 + cb : sb0 + 7 * ROWB + cb;
        const unsigned e1 = dl ? sb2 + cb : sb1 + 7 * ROWB + cb;
        float h[P];
        #pragma unroll
        for (int q = 0; q < P; q++) h[q] = __ldg(prm.taps + q * M + p0);
        float2 acc[kR];
        #pragma unroll
        for (int r = 0; r < kR; r++) acc[r] = make_float2(0.f, 0.f);
        #pragma unroll
        for (int ii = 0; ii < P + kR - 1; ii++) {
          const int j = ii + J0;
          const unsigned addr = (j & 7) == 7 ? (j < 8 ? e0 : e1) : ((j < 8 ? r0b : (j < 16 ? r1b : r2b)) + (j & 7) * ROWB);
          uint32_t w;
          if (IN16) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr));
          else asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w) : "r"(addr));
          const float2 x = unpack<IN16, UNPACK>(w);
          #pragma unroll
          for (int r = 0; r < kR; r++) {
            const int q = r + P - 1 - ii;
            if (q >= 0 && q < P) acc[r] = __ffma2_rn(make_float2(h[q], h[q]), x, acc[r]);
          }
        }
        const int pos = tpad((p0 - (ph ? D : 0)) & (M - 1));
        #pragma unroll
        for (int r = 0; r < kR; r++) tile[r * TS + pos] = acc[r];
      }
      __syncthreads();                               // the tile is complete, nobody reads the ring any more
      if (ph == os - 1) pending = (k + 1 < k1) ? load_slot(a0 + kR, s_old) : false;

      const int row = t >> 7, tg = t & 127;          // from here on a row belongs to the four warps t >> 7
      {   // pass 0: z_{k0}[j] = W_M^{j k0} sum_q u[j + 128 q] W_8^{q k0}  ->  position 128 k0 + j
        float2 tw[7];
        #pragma unroll
        for (int q = 1; q < 8; q++) tw[q - 1] = __ldg(rp.twn + ((tg * q) & (M - 1)));
        float2* rp0 = tile + row * TS + tg;
        float2 v[8];
        #pragma unroll
        for (int q = 0; q < 8; q++) v[q] = rp0[q * 130];
        dft8(v);
        #pragma unroll
        for (int q = 1; q < 8; q++) v[q] = cmul(v[q], tw[q - 1]);
        #pragma unroll
        for (int q = 0; q < 8; q++) rp0[q * 130] = v[q];
      }
      asm volatile("bar.sync %0, 128;" ::"r"(row + 1) : "memory");
      {   // pass 1 inside block k0
        const int j1 = tg & 15, kb = tg >> 4;
        float2* rp1 = tile + row * TS + kb * 130 + j1;
        float2 v[8], tw[7];
        #pragma unroll
        for (int q = 0; q < 8; q++) v[q] = rp1[q * 16];
        #pragma unroll
        for (int q = 1; q < 8; q++) tw[q - 1] = tw1s[(q - 1) * 16 + j1];
        dft8(v);
        #pragma unroll
        for (int q = 1; q < 8; q++) v[q] = cmul(v[q], tw[q - 1]);
        #pragma unroll
        for (int q = 0; q < 8; q++) rp1[q * 16] = v[q];
      }
      asm volatile("bar.sync %0, 128;" ::"r"(row + 1) : "memory");
      if (tg < 64) {   // pass 2: the lower two warps of the row; the upper two go on to the next phase's FIR
        const int b = tg, kb = b & 7, kc = b >> 3;
        const float4* src = (const float4*)(tile + row * TS + kb * 130 + kc * 16);
        float2 v[16];
        #pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 f = src[q];
          v[2 * q] = make_float2(f.x, f.y); v[2 * q + 1] = make_float2(f.z, f.w);
        }
        dft16(v);
        const long long m = (a0 + row) * os + ph;
        if (m >= prm.row_base && m < prm.row_base + prm.nrows) {
          float2* g = prm.out + (m - prm.row_base) * (long long)M + b;
          #pragma unroll
          for (int q = 0; q < 16; q++) g[q * 64] = v[q];
        }
      }
    }
    s_old = s_mid;
  }
}

}  // namespace ring
}  // namespace chzi
