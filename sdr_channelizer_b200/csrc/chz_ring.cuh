// Fused K1+K2+K3 for large channel counts (M = 1024 per CTA): one global read of the raw samples, one
// global write of the channel rows -- the path the split FIR + FFT kernels cannot reach (they move the
// fp32 FIR output to DRAM and back: 28 B per sample against 12 algorithmic).
// Reference math replaced: matlab/create_pdws_channelized.m:35-38 (normalise) and :57 (channelizer(iq)).
//
// One persistent CTA per SM owns a contiguous run of the recording.  On chip it keeps
//   * a ring of the last 24 raw frames A_a = x[aM, aM+M) (int16 pairs as recorded, 4 KB each), fed by
//     cp.async.bulk (TMA, mbarrier completion) eight frames at a time, prefetched into L2 one step ahead;
//   * the P taps of four polyphase branches per FIR thread in registers;
//   * two tiles of 8 output rows x M channels (fp32 complex): the FIR warps fill one while the FFT warps
//     transform the other IN PLACE.
// Per step a FIR thread reads the P+7 raw words of each of its branches once from the ring, unpacks them and feeds
// eight running sums per branch (no re-reads of the recording from L2: DRAM and L2 see every sample once);
// the FFT warps run a decimation-in-frequency FFT 8 x 8 x 16 whose last pass streams the rows to global memory
// with the lanes across channels (whole 256-byte runs per store).  2x oversampling = the same step run twice
// per ring advance: the odd rows read the ring half a frame later and rotate the branches by M/2.
#pragma once
#include "chz_kernels.cuh"

namespace chzi {
namespace ring {

constexpr int kM = 1024;        // channels = branches per CTA
constexpr int kNT = 512;        // threads: 8 FIR warps (four branches per thread) + 8 FFT warps
constexpr int kR = 8;           // output rows per tile = frames per ring slot
constexpr int kSlots = 3;       // ring = 3 slots of 8 frames: rows a0-16 .. a0+7 of the step at a0
constexpr int kTileStride = kM + 2 * (kM / 128) - 2;   // float2 per tile row: 2 pad elements after each 128-block but the last (see pass 2)

struct RingParams {
  long long a_lo;       // first frame whose rows this launch produces (row m = os*a + phase)
  long long nsteps;     // steps of 8 frames over all CTAs
  const float2* twn;    // e^{+j 2 pi i / M}, i < M
  int dbg;              // timing experiments only (CHZ_RING_DBG), bit flags: 1 skip the FFT passes, 2 skip the FIR, 4 skip the ring loads; results are garbage
};

template <bool IN16> struct Smem {
  typedef typename RawT<IN16>::type raw_t;
  static constexpr int ROWB = kM * (int)sizeof(raw_t);            // bytes per frame
  static constexpr int SLOTB = kR * ROWB;                         // bytes per ring slot
  static constexpr int RING = kSlots * SLOTB;
  static constexpr int TILE = 2 * kR * kTileStride * (int)sizeof(float2);   // two tile buffers: the FIR of a phase writes one while the FFT of the previous phase still reads the other
  static constexpr int OFF_TILE = RING;
  static constexpr int OFF_H0 = OFF_TILE + TILE;                  // taps of branch 0 (32 floats)
  static constexpr int OFF_TW1 = OFF_H0 + 32 * (int)sizeof(float);   // pass-1 twiddles W_128^{j1 k}: [7][16] float2
  static constexpr int OFF_BAR = OFF_TW1 + 7 * 16 * (int)sizeof(float2);
  static constexpr int TOTAL = OFF_BAR + 16;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// Unpack policies for one raw word (int16 I | int16 Q << 16, or int8 pair) -> integer-valued float2.
//   0: two I2F.S16 (conversion pipe behind the MIO queue)
//   1: I2F.S16 for I, arithmetic shift + I2FP.F32.S32 (ALU pipe) for Q  -- default, 4 % faster than 0 on B200
// Also measured and dropped (profiles/r02b_ring_unpack_ab.jsonl): PRMT sign extension + I2FP for both halves, and the
// magic-number form ((x ^ 0x8000) | 0x4B400000, one packed FADD2): both within 1 % of policy 1 -- the FIR is bound by
// register-file bandwidth of its FFMA2s (2.8 cycles each with three distinct operands, tools/ubench/pipes.cu), not by
// the conversions.
template <bool IN16, int UNPACK>
__device__ __forceinline__ float2 unpack(uint32_t raw) {
  if (!IN16 || UNPACK == 0) return unpack_raw<IN16>(raw);
  return make_float2((float)(short)(raw & 0xffffu), __int2float_rn(((int)raw) >> 16));
}

// tile element index of channel position pos: two pad elements per 128 so that the eight lanes of a
// quarter-warp that read 16-byte pieces of eight different 128-blocks in pass 2 hit distinct banks
__device__ __forceinline__ int tpad(int pos) { return pos + ((pos >> 7) << 1); }

// ---- the kernel: FIR warps and FFT warps over the two tile buffers -------------------------------------------------
// A kernel in which every warp alternates between the FIR (FMA pipe / register-file bandwidth bound) and the FFT
// passes (shared-memory bound) leaves each pipe idle while the other works: that was the first round-2 kernel
// (k_chan_ring in chz_ring_exp.cuh: 273 / 136 GS/s critical / 2x).  Here warps 0-7 only filter (tile n+1) while
// warps 8-15 only transform (tile n): a FIR thread carries the taps of FOUR branches (two column pairs, filtered one
// after the other with the same 16 accumulators) and raises its register budget with setmaxnreg, an FFT thread takes
// four radix-8 butterflies per pass (two at a time) and lowers it.  Hand-off per tile buffer through named barriers in
// the producer/consumer pattern (bar.arrive by one role, bar.sync by the other); the ring and its TMA copies belong to
// the FIR warps alone.  Same arithmetic per row as k_chan_ring: bit-identical results.
// Measured (profiles/r02j_ring_ws_ab.jsonl): 310 / 162 GS/s critical / 2x at P = 16 (57.5 % / 50.1 % of HBM),
// 340 GS/s at P = 12, 362 GS/s at P = 8.
//
// Register split between the roles (setmaxnreg; the sum over 8 + 8 warps must stay within 65 536): the FIR code needs
// 134 registers at P = 16.  A/B over 128/128, 136/120, 144/112, 152/104, 160/96: all within +-3 % (instruction
// scheduling differences); 136/120 is best at P = 16 and 8, 152/104 at P = 12.
// FW = 16 FFT warps (768 threads: four row groups of two rows, one butterfly in flight per thread, 128/56 registers)
// was built to give the FFT role more warps to hide its latencies with: 267 / 142 GS/s against 310 / 162 with FW = 8
// (EXPERIMENTS build, CHZ_RING_VARIANT=3) -- at 56 registers the role spills and loses its two-butterfly batches.
template <int P, int FW> struct RoleRegs {       // FW = FFT warps (8 or 16) next to the 8 FIR warps
  static constexpr int FIR = FW == 16 ? 128 : (P == 12 ? 152 : 136);
  static constexpr int FFT = FW == 16 ? 56 : (P == 12 ? 104 : 120);
  // setmaxnreg.inc only gets what setmaxnreg.dec of the SAME block released: the budget is what the block was launched
  // with (registers per thread rounded down to 8), not the whole register file -- asking for more waits for ever
  static constexpr int LAUNCH = 65536 / (256 + 32 * FW) / 8 * 8;
  static_assert(8 * 32 * FIR + FW * 32 * FFT <= (256 + 32 * FW) * LAUNCH, "register budget of the block");
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int P, bool IN16, int UNPACK, int FW = 8>
__global__ void __launch_bounds__(256 + 32 * FW, 1) k_chan_ring_ws(ChanParams prm, RingParams rp) {
  constexpr int NTH = 256 + 32 * FW;                 // threads of the block: every hand-off barrier counts them all
  constexpr int NG = FW / 4, RG = kR / NG;           // FFT row groups of 128 threads, rows per group
  typedef Smem<IN16> SM;
  typedef typename RawT<IN16>::type raw_t;
  constexpr int M = kM, ROWB = SM::ROWB, SLOTB = SM::SLOTB, TS = kTileStride;
  constexpr int J0 = 16 - P;
  constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_FIR = 5, BAR_FFT = 6;   // named barriers: full[2], empty[2], FIR group, FFT row groups [2]
  extern __shared__ __align__(128) unsigned char smem[];
  float2* tiles = (float2*)(smem + SM::OFF_TILE);
  float* h0s = (float*)(smem + SM::OFF_H0);
  float2* tw1s = (float2*)(smem + SM::OFF_TW1);
  const unsigned ring_s = smem_u32(smem), bar = smem_u32(smem + SM::OFF_BAR);
  const int t = threadIdx.x;
  const int os = prm.os, D = prm.D;
  const long long k0 = rp.nsteps * blockIdx.x / gridDim.x, k1 = rp.nsteps * (blockIdx.x + 1) / gridDim.x;
  if (k0 >= k1) return;
  if (t < P) h0s[t] = __ldg(prm.taps + t * M);
  if (t < 7 * 16) tw1s[t] = __ldg(rp.twn + ((((t & 15) * ((t >> 4) + 1)) << 3) & (M - 1)));
  if (t == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long nphases = (k1 - k0) * os;

  if (t < 256) {
    // =========================== FIR warps ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RoleRegs<P, FW>::FIR));
    const int f = t;
    float hA1[P], hA2[P], hB1[P], hB2[P];           // taps of the column pairs u = f and u = f + 256
    {
      const int a1 = 2 * f + 1, a2 = 2 * f + 2, b1 = 2 * (f + 256) + 1, b2 = (2 * (f + 256) + 2) & (M - 1);
      #pragma unroll
      for (int q = 0; q < P; q++) {
        hA1[q] = __ldg(prm.taps + q * M + a1); hA2[q] = __ldg(prm.taps + q * M + a2);
        hB1[q] = __ldg(prm.taps + q * M + b1); hB2[q] = __ldg(prm.taps + q * M + b2);
      }
    }
    const long long in_end = prm.in_base + prm.n_in;
    const bool aligned = ((((unsigned long long)prm.in) - (unsigned long long)prm.in_base * sizeof(raw_t)) & 15ull) == 0;
    const raw_t* __restrict__ inp = (const raw_t*)prm.in - prm.in_base;
    unsigned parity = 0;
    bool pending = false;
    auto load_slot = [&](long long a, int s) -> bool {      // FIR warps only
      const long long lo = a * M, hi = lo + (long long)kR * M;
      if (aligned && lo >= prm.in_base && hi <= in_end && !(rp.dbg & 4)) {
        if (f == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(bar, SLOTB);
          bulk_g2s(ring_s + s * SLOTB, inp + lo, SLOTB, bar);
        }
        return true;
      }
      raw_t* dst = (raw_t*)(smem + s * SLOTB);
      if (!(rp.dbg & 4)) {
        #pragma unroll 4
        for (int e = f; e < kR * M; e += 256) dst[e] = (raw_t)load_raw<IN16>(prm, lo + e);
      }
      return false;
    };
    // the copy of step k+1's frames can only be issued when step k is through (all three slots are live until then);
    // asking L2 for them a step ahead takes the HBM latency out of that copy (+4 % measured; two steps ahead: no
    // further gain; refilling the slot in two column halves as each falls dead: -5 %, the extra barrier and the
    // sixteen 2 KB copies cost more than the wait they remove)
    auto prefetch_slot = [&](long long a) {
      const long long lo = a * M, hi = lo + (long long)kR * M;
      if (f == 0 && aligned && lo >= prm.in_base && hi <= in_end)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(inp + lo), "r"(SLOTB) : "memory");
    };
    const long long a_start = rp.a_lo + k0 * kR;
    if (load_slot(a_start - 2 * kR, 0)) { mbar_wait(bar, parity); parity ^= 1; }
    bar_sync(BAR_FIR, 256);
    if (load_slot(a_start - kR, 1)) { mbar_wait(bar, parity); parity ^= 1; }
    bar_sync(BAR_FIR, 256);
    pending = load_slot(a_start, 2);

    int s_old = 0;
    long long n = 0;                                 // phase counter: tile buffer n & 1
    for (long long k = k0; k < k1; k++) {
      const int s_mid = s_old == 2 ? 0 : s_old + 1, s_new = s_mid == 2 ? 0 : s_mid + 1;
      if (pending) { mbar_wait(bar, parity); parity ^= 1; }
      else bar_sync(BAR_FIR, 256);                   // bounds-checked copy by the FIR threads: visible after their barrier
      if (k + 1 < k1) prefetch_slot(rp.a_lo + (k + 1) * kR);
      const unsigned sb0 = ring_s + s_old * SLOTB, sb1 = ring_s + s_mid * SLOTB, sb2 = ring_s + s_new * SLOTB;
      for (int ph = 0; ph < os; ph++, n++) {
        const int buf = (int)(n & 1);
        float2* tile = tiles + buf * (kR * TS);
        if (n >= 2) bar_sync(BAR_EMPTY + buf, NTH);  // the FFT warps have drained this buffer (tile n - 2)
        const int shift = ph ? D : 0;
        // one column pair: 8 rows x branches (2u+2, 2u+1)
        auto fir_pair = [&](int u, const float (&h1)[P], const float (&h2)[P]) {
          const bool lowhalf = ph && u < D / 2;
          const int cl = ph ? (lowhalf ? D - 2 * u - 2 : M + D - 2 * u - 2) : M - 2 * u - 2;
          const unsigned dcol = (unsigned)cl * sizeof(raw_t) + (lowhalf ? ROWB : 0);
          const unsigned r0b = sb0 + dcol, r1b = sb1 + dcol, r2b = sb2 + dcol;
          const unsigned e0 = lowhalf ? sb1 + cl * (unsigned)sizeof(raw_t) : sb0 + 7 * ROWB + cl * (unsigned)sizeof(raw_t);
          const unsigned e1 = lowhalf ? sb2 + cl * (unsigned)sizeof(raw_t) : sb1 + 7 * ROWB + cl * (unsigned)sizeof(raw_t);
          float2 acc1[kR], acc2[kR];
          #pragma unroll
          for (int r = 0; r < kR; r++) { acc1[r] = make_float2(0.f, 0.f); acc2[r] = make_float2(0.f, 0.f); }
          #pragma unroll
          for (int ii = 0; ii < P + kR - 1; ii++) {
            const int j = ii + J0;
            const unsigned addr = (j & 7) == 7 ? (j < 8 ? e0 : e1) : ((j < 8 ? r0b : (j < 16 ? r1b : r2b)) + (j & 7) * ROWB);
            uint32_t wa, wb;
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
              if (q >= 0 && q < P) acc1[r] = __ffma2_rn(make_float2(h1[q], h1[q]), xb, acc1[r]);
            }
          }
          // A warp storing one branch per lane writes every other element (8 bytes at a 16-byte stride: four
          // shared-memory wavefronts for 256 bytes).  Every other group of 8 lanes stores its branches in the opposite order, so that
          // each instruction covers all 16 bank pairs twice: two wavefronts.
          const int pos1 = tpad((2 * u + 1 - shift) & (M - 1)), pos2 = tpad((2 * u + 2 - shift) & (M - 1));
          const bool swp = (u & 8) != 0;
          float2* t1 = tile + (swp ? pos2 : pos1);
          float2* t2 = tile + (swp ? pos1 : pos2);
          #pragma unroll
          for (int r = 0; r < kR; r++) {
            t1[r * TS] = swp ? acc2[r] : acc1[r];
            t2[r * TS] = swp ? acc1[r] : acc2[r];
          }
        };
        auto fix_branch0 = [&]() {
          if (f < 224) return;
          // branch 0 (second element of pair u = 511, written just above by lane 31 of this warp): its frame offset
          // differs from its neighbour's, so eight lanes recompute it from frame (a - q), column ph*D
          __syncwarp();
          const int r = f & 31;
          if (r < kR) {
            float2 acc = make_float2(0.f, 0.f);
            const unsigned c0 = (unsigned)(ph ? D : 0) * sizeof(raw_t);
            #pragma unroll
            for (int q = P - 1; q >= 0; q--) {
              const int i = 16 + r - q;
              const unsigned sb = i < 8 ? sb0 : (i < 16 ? sb1 : sb2);
              uint32_t w;
              if (IN16) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(sb + (i & 7) * ROWB + c0));
              else asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w) : "r"(sb + (i & 7) * ROWB + c0));
              acc = __ffma2_rn(make_float2(h0s[q], h0s[q]), unpack<IN16, UNPACK>(w), acc);
            }
            tile[r * TS + tpad((0 - shift) & (M - 1))] = acc;
          }
        };
        if (!(rp.dbg & 2)) {
          fir_pair(f, hA1, hA2);
          fir_pair(f + 256, hB1, hB2);
          fix_branch0();
        }
        bar_arrive(BAR_FULL + buf, NTH);             // tile n is complete
      }
      // the oldest ring slot is dead once every FIR warp is through this step: request the next step's frames
      bar_sync(BAR_FIR, 256);
      pending = k + 1 < k1 ? load_slot(rp.a_lo + (k + 1) * kR, s_old) : false;
      s_old = s_mid;
    }
  } else {
    // =========================== FFT warps ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RoleRegs<P, FW>::FFT));
    const int g = t - 256;
    const int rr = g >> 7, tg = g & 127;             // rows rr, rr + NG, rr + 2 NG, ... belong to the 4 warps g >> 7
    constexpr int BT = FW == 8 ? 2 : 1;              // radix-8 butterflies in flight per thread (registers)
    float2 tw0[7];
    #pragma unroll
    for (int q = 1; q < 8; q++) tw0[q - 1] = __ldg(rp.twn + ((tg * q) & (M - 1)));
    for (long long n = 0; n < nphases; n++) {
      const int buf = (int)(n & 1);
      float2* tile = tiles + buf * (kR * TS);
      const long long a0 = rp.a_lo + (k0 + n / os) * kR;
      const int ph = (int)(n % os);
      bar_sync(BAR_FULL + buf, NTH);                 // the FIR warps have finished tile n
      if (rp.dbg & 1) { bar_arrive(BAR_EMPTY + buf, NTH); continue; }
      #pragma unroll
      for (int i0 = 0; i0 < RG; i0 += BT) {          // pass 0: one radix-8 butterfly per row of the group
        float2 v[BT][8];
        #pragma unroll
        for (int bi = 0; bi < BT; bi++) {
          const float2* row = tile + (rr + NG * (i0 + bi)) * TS + tg;
          #pragma unroll
          for (int q = 0; q < 8; q++) v[bi][q] = row[q * 130];
        }
        #pragma unroll
        for (int bi = 0; bi < BT; bi++) {
          dft8(v[bi]);
          #pragma unroll
          for (int q = 1; q < 8; q++) v[bi][q] = cmul(v[bi][q], tw0[q - 1]);
        }
        #pragma unroll
        for (int bi = 0; bi < BT; bi++) {
          float2* row = tile + (rr + NG * (i0 + bi)) * TS + tg;
          #pragma unroll
          for (int q = 0; q < 8; q++) row[q * 130] = v[bi][q];
        }
      }
      bar_sync(BAR_FFT + rr, 128);
      {
        const int j1 = tg & 15, kb = tg >> 4;
        float2 tw[7];
        #pragma unroll
        for (int q = 1; q < 8; q++) tw[q - 1] = tw1s[(q - 1) * 16 + j1];
        #pragma unroll
        for (int i0 = 0; i0 < RG; i0 += BT) {        // pass 1
          float2 v[BT][8];
          #pragma unroll
          for (int bi = 0; bi < BT; bi++) {
            const float2* row = tile + (rr + NG * (i0 + bi)) * TS + kb * 130 + j1;
            #pragma unroll
            for (int q = 0; q < 8; q++) v[bi][q] = row[q * 16];
          }
          #pragma unroll
          for (int bi = 0; bi < BT; bi++) {
            dft8(v[bi]);
            #pragma unroll
            for (int q = 1; q < 8; q++) v[bi][q] = cmul(v[bi][q], tw[q - 1]);
          }
          #pragma unroll
          for (int bi = 0; bi < BT; bi++) {
            float2* row = tile + (rr + NG * (i0 + bi)) * TS + kb * 130 + j1;
            #pragma unroll
            for (int q = 0; q < 8; q++) row[q * 16] = v[bi][q];
          }
        }
      }
      bar_sync(BAR_FFT + rr, 128);
      #pragma unroll 1
      for (int hh = 0; hh < RG / 2; hh++) {          // pass 2: one radix-16 butterfly per pair of rows of the group
        const int row_i = rr + NG * ((tg >> 6) & 1) + 2 * NG * hh, b = tg & 63;
        const int kb = b & 7, kc = b >> 3;
        const float4* src = (const float4*)(tile + row_i * TS + kb * 130 + kc * 16);
        float2 v[16];
        #pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 f4 = src[q];
          v[2 * q] = make_float2(f4.x, f4.y); v[2 * q + 1] = make_float2(f4.z, f4.w);
        }
        if (hh == RG / 2 - 1) bar_arrive(BAR_EMPTY + buf, NTH);   // this thread's last read of the tile is in registers
        dft16(v);
        const long long m = (a0 + row_i) * os + ph;
        if (m >= prm.row_base && m < prm.row_base + prm.nrows) {
          float2* gp = prm.out + (m - prm.row_base) * (long long)M + b;
          #pragma unroll
          for (int q = 0; q < 16; q++) gp[q * 64] = v[q];
        }
      }
    }
  }
}


}  // namespace ring
}  // namespace chzi
