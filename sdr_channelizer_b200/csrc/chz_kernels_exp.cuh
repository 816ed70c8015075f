// Experimental large-M / warp-specialised kernels (CHZ_OPT_FORCE_PATH 3..10).  Each lost its A/B against the
// default paths; they are built only with `make EXPERIMENTS=1`.
#pragma once
#include "chz_kernels.cuh"

namespace chzi {

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
      const int r = padi_first<M>((p - sp.shift + M) % M);
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
    const unsigned pos = (unsigned)padi_first<M>((p - sp.shift + M) % M) * (unsigned)sizeof(float2);
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
