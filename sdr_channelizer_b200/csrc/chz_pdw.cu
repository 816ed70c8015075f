// K4: channelized PDW extraction on the GPU.
// Replaces matlab/create_pdws_channelized.m:60-136:
//   :60      fftshift            -> channel index remap when records are built (no data movement)
//   :67      mag = abs(iq)       -> mag_of() recomputed from the fp32 channel matrix in every pass
//   :73      median(mag)         -> exact per-channel radix select (3 histogram passes, 11+11+9 bits, on mag^2)
//   :75      threshold           -> host, double precision, then bracketed by two floats
//   :79-96   edge FSM            -> k_detect: lanes = channels (coalesced), rows in lock-step, warp
//                                   ballot + one atomic per warp to compact edge events
//   :97-128  per-pulse stats     -> k_pulse_stats: one block per pulse, radix-select medians
// Input layout: y[row][k], natural channel order, fp32 complex.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "chz_internal.h"
#include "chz_launch.h"

namespace chzi {

__device__ __forceinline__ float mag2_of(float2 v) { return __fmaf_rn(v.x, v.x, __fmul_rn(v.y, v.y)); }
__device__ __forceinline__ float mag_of(float2 v) {   // |y| in fp32, one rounding sequence everywhere
  return __fsqrt_rn(mag2_of(v));
}

// ---- exact per-channel median: radix select over the bit pattern of non-negative floats ------------
constexpr int kBins = 2048;
struct SelState {          // per channel, two order statistics (lower and upper middle element)
  uint32_t prefix[2];      // bits fixed so far (high bits)
  uint32_t rank[2];        // rank still to find inside the prefix bucket
};

// pass 0: bits 30..20, pass 1: bits 19..9, pass 2: bits 8..0
__device__ __forceinline__ int pass_shift(int pass) { return pass == 0 ? 20 : (pass == 1 ? 9 : 0); }
__device__ __forceinline__ uint32_t pass_mask(int pass) { return pass == 2 ? 0x1FFu : 0x7FFu; }
__device__ __forceinline__ uint32_t prefix_mask(int pass) { return pass == 0 ? 0u : (pass == 1 ? 0xFFF00000u : 0xFFFFFE00u); }

// ge = smallest float >= the leading threshold, le = largest float <= the trailing one; ge2 / le2 = the same two
// decisions on the SQUARED magnitude: |y| >= ge <=> |y|^2 >= ge2 and |y| <= le <=> |y|^2 <= le2 for every float |y|^2,
// because |y| = sqrt_rn(|y|^2) is monotone (squared_bounds walks to the exact boundaries).  The detector never takes
// a square root.
struct Thr { float ge, le, ge2, le2; };
__host__ __device__ inline float sqrt_rn_(float x) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(x);
#else
  return sqrtf(x);                                       // IEEE: correctly rounded, like sqrt.rn.f32
#endif
}
__host__ __device__ inline float bits_step_(float x, int d) {   // next (d = +1) / previous (d = -1) float of a non-negative x
  uint32_t b;
  memcpy(&b, &x, 4);
  b += (uint32_t)d;
  memcpy(&x, &b, 4);
  return x;
}
__host__ __device__ inline void squared_bounds(Thr* t) {
  const float ge = t->ge, le = t->le, big = 3.402823466e38f;
  // ge2 = the smallest x >= 0 with sqrt_rn(x) >= ge
  if (ge != ge) t->ge2 = ge;                             // NaN: never true, like the comparison with ge itself
  else if (!(ge > 0.f)) t->ge2 = 0.f;                    // every magnitude qualifies
  else if (ge > big) t->ge2 = ge;                        // +inf
  else {
    float x = ge * ge;
    if (x > big) x = big;
    for (int i = 0; i < 64 && x > 0.f && sqrt_rn_(x) >= ge; i++) x = bits_step_(x, -1);
    for (int i = 0; i < 64 && sqrt_rn_(x) < ge; i++) x = bits_step_(x, +1);      // steps from FLT_MAX to +inf if need be
    t->ge2 = x;
  }
  // le2 = the largest x with sqrt_rn(x) <= le (none if le < 0)
  if (le != le) t->le2 = le;
  else if (le < 0.f) t->le2 = -1.f;
  else if (le > big) t->le2 = le;
  else {
    float x = le * le;
    if (x > big) x = big;
    for (int i = 0; i < 64 && x <= big && sqrt_rn_(x) <= le; i++) x = bits_step_(x, +1);
    for (int i = 0; i < 64 && x > 0.f && sqrt_rn_(x) > le; i++) x = bits_step_(x, -1);
    t->le2 = x;
  }
}
__device__ __forceinline__ void thresholds_of(const SelState& s, double scale, double scale_lo, Thr* thr, double* nf);

// One 256-thread block and one channel: find the bucket holding each wanted rank, fix its bits, reduce the rank, clear
// the channel's histogram rows for the next pass.  All threads of the block call it.
// thr != NULL (last pass of the one-GPU extractor): the noise floor and the thresholds are derived right here,
// which saves the k_thresholds launch.
struct SelArgs { uint32_t rank_lo, rank_hi; double scale, scale_lo; Thr* thr; double* nf; };
__device__ __forceinline__ void select_channel(uint32_t* __restrict__ hist, SelState* __restrict__ st, int pass, int ch,
                                               const SelArgs& sa, uint32_t* scratch /* shared, 2 * 256 + 4 words */) {
  uint32_t* const res_bin = scratch + 512;
  uint32_t* const res_rank = scratch + 514;
  const uint32_t rank_lo = sa.rank_lo, rank_hi = sa.rank_hi;
  const double scale = sa.scale, scale_lo = sa.scale_lo;
  Thr* const thr = sa.thr;
  double* const nf = sa.nf;
  SelState s;
  if (pass == 0) { s.prefix[0] = s.prefix[1] = 0; s.rank[0] = rank_lo; s.rank[1] = rank_hi; }
  else s = st[ch];
  const bool split = pass > 0 && s.prefix[0] != s.prefix[1];
  const int nb = pass == 2 ? 512 : kBins, per = nb / 256;
  // 256 partial sums per slot, then warp `slot` finds the bucket that holds the wanted rank with a prefix sum
  // across its lanes (one thread walking 256 partial sums took 10-15 us per pass)
  uint32_t (*part2)[256] = reinterpret_cast<uint32_t (*)[256]>(scratch);
  for (int slot = 0; slot < 2; slot++) {
    const uint32_t* hrow = hist + ((size_t)ch * 2 + ((slot == 1 && split) ? 1 : 0)) * kBins;
    uint32_t loc = 0;
    for (int i = 0; i < per; i++) loc += hrow[threadIdx.x * per + i];
    part2[slot][threadIdx.x] = loc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 2) {
    const int slot = warp;
    const uint32_t* hrow = hist + ((size_t)ch * 2 + ((slot == 1 && split) ? 1 : 0)) * kBins;
    const uint32_t want = s.rank[slot];
    uint32_t c[8], sum = 0;
    #pragma unroll
    for (int j = 0; j < 8; j++) { c[j] = part2[slot][lane * 8 + j]; sum += c[j]; }
    uint32_t incl = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const uint32_t excl = incl - sum;
    const bool here = excl <= want && want < incl;
    const unsigned any = __ballot_sync(0xffffffffu, here);
    if (here || (any == 0 && lane == 31)) {          // no lane qualifies only if want >= total: clamp to the last bucket
      uint32_t acc = excl;
      int t = lane * 8;
      #pragma unroll
      for (int j = 0; j < 7; j++) {
        if (t == lane * 8 + j) { if (acc + c[j] > want) { /* found */ } else { acc += c[j]; t++; } }
      }
      int b = t * per;
      for (; b < t * per + per - 1; b++) { if (acc + hrow[b] > want) break; acc += hrow[b]; }
      res_bin[slot] = (uint32_t)b;
      res_rank[slot] = want - acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int shift = pass == 0 ? 20 : (pass == 1 ? 9 : 0);
    for (int slot = 0; slot < 2; slot++) { s.prefix[slot] |= res_bin[slot] << shift; s.rank[slot] = res_rank[slot]; }
    st[ch] = s;
    if (thr) thresholds_of(s, scale, scale_lo, thr + ch, nf + ch);
  }
  __syncthreads();
  // clear both histogram rows of this channel for the next pass
  for (int i = threadIdx.x; i < 2 * kBins; i += 256) hist[(size_t)ch * 2 * kBins + i] = 0;
  __syncthreads();                                       // callers loop over channels: the shared scratch is reused
}
// one block per channel (time-sharded extraction, where the host sums the shards' histograms between the passes)
__global__ void __launch_bounds__(256) k_select(uint32_t* __restrict__ hist, SelState* __restrict__ st, int pass, SelArgs sa) {
  __shared__ uint32_t scratch[2 * 256 + 4];
  select_channel(hist, st, pass, blockIdx.x, sa, scratch);
}


// Local histogram of one radix pass of the per-channel median select.
// grid: (slabs of 16 channels, row chunks); block 256 = 16 lanes across channels x 16 rows per step, so a warp's load is
// two whole 128-byte lines (16 channels x 2 rows).  The first version gave a block FOUR channels (32-byte row segments,
// 32-bit counters): eight sectors of eight different rows per warp load, 1.7 TB/s on a cold 45 MB matrix and
// 2.3 TB/s on a 2 GB one; this one runs the 2 GB matrix at 5.8 TB/s (340 us per pass, 90 % of the copy peak).
// Sixteen histograms of 2048 bins fit a block as 16-bit counters packed in pairs (64 KB, three blocks per SM); the host
// keeps a block's row range below 65 536 so that no counter can overflow.  Channel c's counters start at word
// c * 1025: the odd stride spreads equal bins of different channels (the rule: the channels' noise floors are alike)
// over the banks.  M < 16 dividing 16: a "row" of 16 lanes is rpw = 16 / M consecutive matrix rows (all lanes busy,
// loads still contiguous); other M < 16 or M % 16 != 0: the lanes past the last channel idle.
// hist: [M][2][kBins] uint32.  Slot 0 is privatised in shared memory; slot 1 (only used when the two order statistics
// have diverged into different buckets) goes straight to global atomics.
// The select runs on |y|^2 (the operand of mag_of's square root): the correctly rounded square root is monotone, so
// the k-th smallest |y| is the square root of the k-th smallest |y|^2 (thresholds_of takes it), and the binning loop
// is a dozen instructions per element (with the square root, 64-bit row checks on every load and a branchy body it
// was 76 and the pass was issue bound).  Whole batches of 8 loads skip the row checks; the last partial batch keeps them.
// Tried and removed: running the select inside this launch, by the block that finishes a channel slab last (a ticket
// per slab): the selects of a slab then run one after the other in ONE block at the tail of the launch (about
// 5 us each: dependent global reads and four barriers) where k_select spreads them over M blocks -- 52 + 65 us for the
// three passes against 34 + 46 us with the separate launches.
constexpr int kH16Stride = kBins / 2 + 1;
constexpr int kH16Smem = 16 * kH16Stride * (int)sizeof(uint32_t);
__global__ void __launch_bounds__(256, 3) k_hist(const float2* __restrict__ y, long long nrows, int M, int pass, int rpw,
                                                 const SelState* __restrict__ st, uint32_t* __restrict__ hist) {
  extern __shared__ __align__(16) uint32_t sh16[];
  for (int i = threadIdx.x; i < 16 * kH16Stride; i += 256) sh16[i] = 0;
  const int lane16 = threadIdx.x & 15;
  const int cl = rpw > 1 ? lane16 % M : lane16, sub = rpw > 1 ? lane16 / M : 0;
  const int ch = blockIdx.x * 16 + cl;
  const bool ch_ok = ch < M;
  const int shift = pass_shift(pass);
  const uint32_t bmask = pass_mask(pass), pmask = prefix_mask(pass);
  uint32_t p0 = 0, p1 = 0;
  if (pass > 0 && ch_ok) { p0 = st[ch].prefix[0]; p1 = st[ch].prefix[1]; }
  const bool split = p0 != p1;
  __syncthreads();
  const long long rows_per_block = (nrows + gridDim.y - 1) / gridDim.y;   // <= 65 535 (host)
  const long long r_begin = (long long)blockIdx.y * rows_per_block;
  long long r_end = r_begin + rows_per_block;
  if (r_end > nrows) r_end = nrows;
  constexpr int UN = 8;
  uint32_t* const shc = sh16 + cl * kH16Stride;
  uint32_t* const g1 = hist + ((size_t)ch * 2 + 1) * kBins;
  auto bin_one = [&](float2 v) {
    const uint32_t bits = __float_as_uint(mag2_of(v));
    const uint32_t pre = bits & pmask, bin = (bits >> shift) & bmask;
    if (pre == p0) atomicAdd(shc + (bin >> 1), (bin & 1u) ? 0x10000u : 1u);
    if (split && pre == p1) atomicAdd(g1 + bin, 1u);
  };
  if (ch_ok) {
    const long long rs = 16LL * rpw;                        // rows per step of the block
    const long long rstep = rs * M;                         // elements between two loads of a thread
    long long r = r_begin + (long long)(threadIdx.x >> 4) * rpw + sub;
    const float2* p = y + r * M + ch;
    for (; r + rs * (UN - 1) < r_end; r += rs * UN, p += UN * rstep) {
      float2 v[UN];
      #pragma unroll
      for (int u = 0; u < UN; u++) v[u] = __ldg(p + u * rstep);
      #pragma unroll
      for (int u = 0; u < UN; u++) bin_one(v[u]);
    }
    if (r < r_end) {                                       // last, partial batch: the same loads in flight, with row checks
      float2 v[UN];
      #pragma unroll
      for (int u = 0; u < UN; u++) v[u] = r + rs * u < r_end ? __ldg(p + u * rstep) : make_float2(0.f, 0.f);
      #pragma unroll
      for (int u = 0; u < UN; u++) if (r + rs * u < r_end) bin_one(v[u]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * (kBins / 2); i += 256) {
    const int c = i / (kBins / 2), w = i % (kBins / 2);
    const uint32_t v = sh16[c * kH16Stride + w];
    if (!v || blockIdx.x * 16 + c >= M) continue;
    uint32_t* g = hist + ((size_t)(blockIdx.x * 16 + c) * 2) * kBins + 2 * w;
    if (v & 0xffffu) atomicAdd(g, v & 0xffffu);
    if (v >> 16) atomicAdd(g + 1, v >> 16);
  }
}

// ---- edge detection -----------------------------------------------------------------------------------
// Leading edge: mag >= T_lead <=> mag >= ge (ge = smallest float >= T_lead).  Trailing edge: mag <= T_trail <=>
// mag <= le (le = largest float <= T_trail).  The channelized script uses one threshold for both
// (create_pdws_channelized.m:88,94: T_trail == T_lead, ge/le bracket it); the wideband script uses
// hysteresis (create_pdws.m:45-47,58,63: 18 dB up, 3 dB down => le < ge).
// Noise floor and thresholds (:73-75) from the selected order statistics, in double like the script, then
// bracketed by floats: ge = smallest float >= T_lead, le = largest float <= T_trail (so the fp32 comparisons
// of k_detect decide exactly like the double comparison would).  On the device so that the extractor does
// not have to stop for a round trip through the host between the median and the edge detector.
__device__ __forceinline__ float float_ge(double t) {
  float f = __double2float_rn(t);
  if ((double)f < t) f = f == 0.f ? __uint_as_float(1u) : (f > 0.f ? __uint_as_float(__float_as_uint(f) + 1u) : __uint_as_float(__float_as_uint(f) - 1u));
  return f;
}
__device__ __forceinline__ float float_le(double t) {
  float f = __double2float_rn(t);
  if ((double)f > t) f = f == 0.f ? __uint_as_float(0x80000001u) : (f > 0.f ? __uint_as_float(__float_as_uint(f) - 1u) : __uint_as_float(__float_as_uint(f) + 1u));
  return f;
}
__device__ __forceinline__ void thresholds_of(const SelState& s, double scale, double scale_lo, Thr* thr, double* nf) {
  // the select ran on |y|^2 (k_hist): the order statistics of |y| are the square roots of those of |y|^2
  const double lo = (double)__fsqrt_rn(__uint_as_float(s.prefix[0])), hi = (double)__fsqrt_rn(__uint_as_float(s.prefix[1]));
  const double v = 0.5 * (lo + hi);                      // MATLAB median: mean of the two middle values
  *nf = v;
  Thr t;
  t.ge = float_ge(v * scale);
  t.le = float_le(v * scale_lo);
  squared_bounds(&t);
  *thr = t;
}
__global__ void k_thresholds(const SelState* __restrict__ st, int M, double scale, double scale_lo, Thr* __restrict__ thr,
                             double* __restrict__ nf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M) return;
  thresholds_of(st[k], scale, scale_lo, thr + k, nf + k);
}

// Lanes walk channels (coalesced 8-byte loads of a row), all lanes advance row by row through the same
// chunk, so a warp ballot per row tells whether any channel saw an edge; one atomic per warp reserves
// the slots and every lane with an edge writes at its ballot rank.
// event = (shifted channel << 40) | (1-based row << 1) | (1 = trailing edge)
//
// PULSES (one-GPU extractor): instead of edge events the kernel emits finished pulses.  A trailing edge knows its
// leading edge when that lay in the same chunk.  Otherwise the pulse goes out with toa = 0 and the statistics kernel
// resolves it from the per-(chunk, channel) summaries written here -- the last row <= le of the chunk and the first
// row >= ge after it -- walking back one 64-row chunk per step (walking back row by row inside this kernel made the
// warps that sit in a long pulse the tail of the launch: 53 us against 17 us for everybody else).  No sort and no
// pairing on the host, so the per-pulse statistics kernel can follow without a round trip.  The rare case the look-back
// cannot decide locally -- a sample exactly equal to a single representable threshold, which toggles the state
// (:88/:94) -- raises count[1] and the host falls back to the event path.
struct PulseIn { unsigned long long toa, end; uint32_t k, kph; };   // rows 1-based, natural channels

template <bool PULSES>
__global__ void __launch_bounds__(256) k_detect(const float2* __restrict__ y, long long nrows, int M,
                                                const Thr* __restrict__ thr, int chunk_rows,
                                                const uint8_t* __restrict__ entry, unsigned long long row_offset,
                                                unsigned long long* __restrict__ events,
                                                unsigned long long cap, unsigned long long* __restrict__ count,
                                                int kph_fixed, uint2* __restrict__ summ) {
  const int lanes_ch = M < 32 ? M : 32;                   // channels per warp
  const int streams = 32 / lanes_ch;                      // independent row chunks inside one warp (M < 32)
  const int lane = threadIdx.x & 31;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int ch_groups = (M + 31) / 32;
  const long long nchunks = (nrows + chunk_rows - 1) / chunk_rows;
  const long long total_warps = nchunks / streams + (nchunks % streams ? 1 : 0);
  const long long wchunk = warp_id / ch_groups;
  if (wchunk >= total_warps) return;                      // warp-uniform
  const int ch_raw = (int)(warp_id % ch_groups) * 32 + (lane % lanes_ch);
  const long long chunk = wchunk * streams + lane / lanes_ch;
  // lanes beyond M (M not a multiple of 32) or beyond streams*lanes_ch idle but still vote in the ballots
  const bool live = chunk < nchunks && ch_raw < M && lane < streams * lanes_ch;
  const int ch = ch_raw < M ? ch_raw : M - 1;
  const long long r0 = chunk * (long long)chunk_rows;
  long long r1 = r0 + chunk_rows;
  if (r1 > nrows) r1 = nrows;
  const Thr t = thr[ch];
  const bool exact = t.ge == t.le;                        // one threshold that is itself a float: equality can occur
  const unsigned long long chs = (unsigned long long)((ch + M / 2) % M);   // fftshift column (:60): k -> (k + floor(M/2)) mod M
  // State on entry = state after row r0-1.  Walking back: a sample >= ge leaves the FSM active, one
  // <= le leaves it inactive whatever came before; samples strictly between the two thresholds
  // (hysteresis only) keep the earlier state; a sample exactly equal to a single representable threshold
  // toggles it (:88 uses >=, :94 uses <= on the same value).  `entry` (time shards, per natural channel) is
  // the state the previous shard left behind; NULL = the FSM starts inactive (:83).
  bool active = entry ? entry[ch] != 0 : false;
  long long lead = -1;                                     // PULSES: 0-based row of the open pulse's leading edge
  if (live && r0 > 0) {
    bool flips = false;
    long long j = r0 - 1;
    for (; j >= 0; j--) {
      const float m2 = mag2_of(y[j * M + ch]);
      const bool a = m2 >= t.ge2, b = m2 <= t.le2;
      if (exact && a && b) { flips = !flips; if (PULSES) count[1] = 1; continue; }
      if (a) { active = true; break; }
      if (b) { active = false; break; }
    }
    active = active != flips;
  }
  constexpr int UN = PULSES ? 16 : 8;                      // rows fetched ahead of the (sequential) state machine
  uint32_t last_below = 0xFFFFFFFFu, first_above = 0xFFFFFFFFu;   // PULSES: chunk summary (rows of this launch)
  // Per batch of UN rows every lane first reduces its rows to two bit masks (bit u: row u is >= ge / <= le; no square
  // root, see Thr) and updates the chunk summary from them with bit scans.  A lane's state can only change in a batch
  // in which an inactive channel sees a row >= ge or an active one a row <= le; if no lane of the warp is in that
  // position -- the rule between pulses -- the batch is done.  Otherwise the warp walks the batch row by row with a
  // ballot per row as before.  (ncu before: 60 warp instructions per element, 31 us per 45 MB.)
  // The next batch's rows are requested before this batch is looked at: with a few dozen warps per SM each walking its
  // chunk batch by batch, the memory latency of every batch was in the open.
  float2 vn[UN];
  #pragma unroll
  for (int u = 0; u < UN; u++) vn[u] = (live && r0 + u < r1) ? __ldg(y + (r0 + u) * M + ch) : make_float2(0.f, 0.f);
  for (int i0 = 0; i0 < chunk_rows; i0 += UN) {            // lock-step over the chunk
    const long long rb = r0 + i0;
    float2 v[UN];
    #pragma unroll
    for (int u = 0; u < UN; u++) v[u] = vn[u];
    if (i0 + UN < chunk_rows) {
      #pragma unroll
      for (int u = 0; u < UN; u++) vn[u] = (live && rb + UN + u < r1) ? __ldg(y + (rb + UN + u) * M + ch) : make_float2(0.f, 0.f);
    }
    uint32_t A = 0, B = 0;
    #pragma unroll
    for (int u = 0; u < UN; u++) {
      const float m2 = mag2_of(v[u]);
      A |= (m2 >= t.ge2 ? 1u : 0u) << u;
      B |= (m2 <= t.le2 ? 1u : 0u) << u;
    }
    const long long left = live ? r1 - rb : 0;            // rows of this batch that exist
    const uint32_t vmask = left >= UN ? (1u << UN) - 1u : (left > 0 ? (1u << (int)left) - 1u : 0u);
    A &= vmask; B &= vmask;
    if (PULSES) {
      if (exact && (A & B)) count[1] = 1;
      if (B) {
        const int hb = 31 - __clz(B);
        last_below = (uint32_t)(rb + hb);
        const uint32_t after = A & ~((2u << hb) - 1u);
        first_above = after ? (uint32_t)(rb + __ffs(after) - 1) : 0xFFFFFFFFu;
      } else if (A && first_above == 0xFFFFFFFFu) {
        first_above = (uint32_t)(rb + __ffs(A) - 1);
      }
    }
    if (!__any_sync(0xffffffffu, active ? B != 0u : A != 0u)) continue;
    #pragma unroll
    for (int u = 0; u < UN; u++) {
      const long long r = rb + u;
      bool ev = false;
      const bool a = (A >> u) & 1u, b = (B >> u) & 1u;
      if (!active) { if (a) { active = true; ev = !PULSES; lead = r; } }   // leading edge (:88)
      else if (b) { active = false; ev = true; }                            // trailing edge (:94)
      const unsigned ball = __ballot_sync(0xffffffffu, ev);
      if (ball) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(ball));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (ev) {
          const unsigned long long slot = base + __popc(ball & ((1u << lane) - 1));
          if (PULSES) {
            if (slot < cap) {
              PulseIn p;
              p.toa = lead >= 0 ? (unsigned long long)(lead + 1) + row_offset : 0ull; p.end = (unsigned long long)(r + 1) + row_offset;
              p.k = (uint32_t)ch; p.kph = kph_fixed >= 0 ? (uint32_t)kph_fixed : (uint32_t)ch;
              *(PulseIn*)((unsigned char*)events + slot * (sizeof(PulseIn) + 24)) = p;   // slot of a {PulseIn, PulseOut} record
            }
          } else if (slot < cap) {
            events[slot] = (chs << 40) | (((unsigned long long)(r + 1) + row_offset) << 1) | (active ? 0ull : 1ull);
          }
        }
      }
    }
  }
  if (PULSES && live) summ[chunk * M + ch] = make_uint2(last_below, first_above);
}

// State a shard leaves behind, per natural channel, as a function of the state it was entered with:
// 0 = inactive, 1 = active (a decisive sample was found walking back from the last row), 2 = the entry
// state, 3 = the entry state toggled (every sample sat between the thresholds or exactly on a single one).
__global__ void __launch_bounds__(128) k_exit_state(const float2* __restrict__ y, long long nrows, int M,
                                                    const Thr* __restrict__ thr, uint8_t* __restrict__ code) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= M) return;
  const Thr t = thr[ch];
  const bool exact = t.ge == t.le;
  bool flips = false;
  int c = 2;
  for (long long j = nrows - 1; j >= 0; j--) {
    const float m2 = mag2_of(y[j * M + ch]);
    const bool a = m2 >= t.ge2, b = m2 <= t.le2;
    if (exact && a && b) { flips = !flips; continue; }
    if (a) { c = 1; break; }
    if (b) { c = 0; break; }
  }
  code[ch] = (uint8_t)(c == 2 ? (flips ? 3 : 2) : (flips ? 1 - c : c));
}

// ---- per-pulse statistics ----------------------------------------------------------------------------
struct PulseOut { float amp_lo, amp_hi, pd_lo, pd_hi; uint32_t sat, pad; };
static_assert(sizeof(PulseOut) == 24 && sizeof(PulseIn) == 24, "k_detect<true> writes PulseIn at a stride of sizeof(PulseIn) + sizeof(PulseOut)");

__device__ __forceinline__ uint32_t fkey(float f) {      // order-preserving map float -> uint32
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
__device__ __forceinline__ double phase_deg(float2 v) {  // rad2deg(angle(iq)) (:68)
  return atan2((double)v.y, (double)v.x) * (180.0 / 3.14159265358979323846264338327950288);
}
__device__ __forceinline__ float wrapped_diff(float2 a, float2 b) {   // :114-116
  double d = phase_deg(b) - phase_deg(a);
  if (d < -180.0) d += 360.0;
  if (d > 180.0) d -= 360.0;
  return (float)d;
}

// Exact k-th smallest (two ranks at once) of n keys produced by key(i), 4 passes of 8 bits.
// The samples of a pulse are nearly equal, so in the high-order passes every thread wants the same bin:
// lanes with the same bin are merged with match.any and ONE of them adds the count (plain shared atomics
// serialised 128 ways), and the bucket holding each rank is found by a warp-wide prefix sum instead of one
// thread walking 256 bins.  blockDim.x must be a multiple of 32 and at least 64.
template <typename KeyF>
__device__ void block_select2(KeyF key, unsigned long long n, unsigned long long rank_lo, unsigned long long rank_hi,
                              uint32_t* h0, uint32_t* h1, uint32_t* out /*[2]*/, unsigned long long* sh_rank) {
  uint32_t pre[2] = {0u, 0u};
  unsigned long long rk[2] = {rank_lo, rank_hi};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long n_up = (n + blockDim.x - 1) / blockDim.x * blockDim.x;   // whole warps stay in the loop
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    const uint32_t pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { h0[i] = 0; h1[i] = 0; }
    __syncthreads();
    const bool split = pre[0] != pre[1];
    for (unsigned long long i = threadIdx.x; i < n_up; i += blockDim.x) {
      int b0 = -1, b1 = -1;
      if (i < n) {
        const uint32_t k = key(i);
        if ((k & pmask) == pre[0]) b0 = (int)((k >> shift) & 0xFF);
        if (split && (k & pmask) == pre[1]) b1 = (int)((k >> shift) & 0xFF);
      }
      const unsigned m0 = __match_any_sync(0xffffffffu, b0);
      if (b0 >= 0 && lane == __ffs(m0) - 1) atomicAdd(&h0[b0], (uint32_t)__popc(m0));
      if (split) {                                           // block-uniform
        const unsigned m1 = __match_any_sync(0xffffffffu, b1);
        if (b1 >= 0 && lane == __ffs(m1) - 1) atomicAdd(&h1[b1], (uint32_t)__popc(m1));
      }
    }
    __syncthreads();
    if (warp < 2) {                                          // warp 0: lower rank, warp 1: upper rank
      const uint32_t* hh = (warp == 1 && split) ? h1 : h0;
      const unsigned long long want = rk[warp];
      uint32_t c[8];
      unsigned long long sum = 0;
      #pragma unroll
      for (int j = 0; j < 8; j++) { c[j] = hh[lane * 8 + j]; sum += c[j]; }
      unsigned long long incl = sum;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const unsigned long long excl = incl - sum;
      const bool here = excl <= want && want < incl;
      const unsigned any = __ballot_sync(0xffffffffu, here);
      if (here || (any == 0 && lane == 31)) {                // (no lane qualifies only if want >= total: clamp to bin 255)
        unsigned long long acc = excl;
        int bsel = lane * 8;
        #pragma unroll
        for (int j = 0; j < 7; j++) {
          if (bsel == lane * 8 + j) { if (acc + c[j] > want) { /* found */ } else { acc += c[j]; bsel++; } }
        }
        out[warp] = (uint32_t)bsel;
        sh_rank[warp] = want - acc;
      }
    }
    __syncthreads();
    for (int s = 0; s < 2; s++) { pre[s] |= out[s] << shift; rk[s] = sh_rank[s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = pre[0]; out[1] = pre[1]; }
  __syncthreads();
}

constexpr int kPulseCap = 5632;   // rows of a pulse held in shared memory (8 B per row: 44 KB)

// npulses_dev != NULL: the pulse count lives on the device (k_detect<true> has just written it) and the blocks
// stride over the list; otherwise one block per pulse of a host-built list (n_host pulses).
template <int STRIDE>   // bytes between consecutive PulseIn (and between consecutive PulseOut) records
__device__ __forceinline__ void pulse_stats_body(const float2* __restrict__ y, long long M, double sat_level,
                                                 unsigned long long row0,
                                                 const PulseIn* __restrict__ in, PulseOut* __restrict__ outp,
                                                 const unsigned long long* __restrict__ npulses_dev,
                                                 unsigned long long n_host, unsigned long long cap,
                                                 const uint2* __restrict__ summ, int chunk_rows) {
  __shared__ unsigned long long sh_toa;
  __shared__ uint32_t h0[256], h1[256], res[2];
  __shared__ unsigned long long sh_rank[2];
  __shared__ int sh_sat;
  __shared__ uint32_t kmag[kPulseCap], kpd[kPulseCap];      // order-preserving keys of |y| and of the phase differences
  unsigned long long total = npulses_dev ? *npulses_dev : n_host;
  if (total > cap) total = cap;
  if (summ && npulses_dev[1]) return;      // k_detect<true> asked for the event path: its pulse list is not to be trusted
  for (unsigned long long pi = blockIdx.x; pi < total; pi += gridDim.x) {
  __syncthreads();                                           // shared state of the previous pulse is no longer read
  PulseIn p = *(const PulseIn*)((const unsigned char*)in + pi * STRIDE);
  if (summ && p.toa == 0) {
    // the pulse was already open when its trailing edge's chunk began: its leading edge is the first row >= ge
    // after the last row <= le before that chunk (k_detect<true>'s summaries, one 64-row chunk per step)
    if (threadIdx.x == 0) {
      uint32_t lead = 0xFFFFFFFFu;
      for (long long c = (long long)((p.end - 1) / chunk_rows) - 1; c >= 0; c--) {
        const uint2 sm = summ[c * M + p.k];
        if (sm.y != 0xFFFFFFFFu) lead = sm.y;
        if (sm.x != 0xFFFFFFFFu) break;
      }
      sh_toa = lead == 0xFFFFFFFFu ? 0ull : (unsigned long long)lead + 1;
      ((PulseIn*)((unsigned char*)in + pi * STRIDE))->toa = sh_toa;
    }
    __syncthreads();
    p.toa = sh_toa;
  }
  if (p.toa == 0 || p.toa > p.end) continue;                 // (never for a consistent list; keeps a bad one from reading out of bounds)
  const unsigned long long a = p.toa - 1 - row0, b = p.end - 1 - row0;   // 0-based inclusive rows of y
  if (threadIdx.x == 0) sh_sat = 0;
  __syncthreads();
  const unsigned long long n1 = b - a + 1, n2 = b - a;
  const bool cached = n1 <= (unsigned long long)kPulseCap;
  // A pulse's column is strided by M*8 bytes, so every pass over it drags whole sectors through DRAM; the
  // selects below make 8 passes.  Pulses of up to kPulseCap rows are therefore read ONCE: the keys of the
  // magnitudes and of the wrapped phase differences and the saturation flag go to shared memory and the
  // selects run there (measured on 8862 pulses over a 1.97 GB matrix: 1.78 GB of DRAM reads and 443 us before).
  int sat = 0;
  if (cached) {
    for (unsigned long long i = threadIdx.x; i < n1; i += blockDim.x) {
      const float2 v = y[(a + i) * M + p.k];
      kmag[i] = fkey(mag_of(v));
      // saturation: rows strictly between the edges (:129-132 runs only while the pulse stays active)
      if (i > 0 && i < n2 && ((double)fabsf(v.x) >= sat_level || (double)fabsf(v.y) >= sat_level)) sat = 1;   // :130
      if (i < n2)                                            // :114-116 (the neighbour's sample comes from L1/L2)
        kpd[i] = fkey(wrapped_diff(p.kph == p.k ? v : y[(a + i) * M + p.kph], y[(a + i + 1) * M + p.kph]));
    }
  } else {
    for (unsigned long long r = a + 1 + threadIdx.x; r < b; r += blockDim.x) {
      const float2 v = y[r * M + p.k];
      if ((double)fabsf(v.x) >= sat_level || (double)fabsf(v.y) >= sat_level) sat = 1;   // :130
    }
  }
  if (sat) atomicOr(&sh_sat, 1);
  __syncthreads();
  PulseOut o;
  // amplitude: median(mag(toa:jj,bin)), both edges included (:101)
  if (cached) block_select2([&](unsigned long long i) { return kmag[i]; }, n1, (n1 - 1) / 2, n1 / 2, h0, h1, res, sh_rank);
  else block_select2([&](unsigned long long i) { return fkey(mag_of(y[(a + i) * M + p.k])); }, n1, (n1 - 1) / 2, n1 / 2,
                     h0, h1, res, sh_rank);
  o.amp_lo = fkey_inv(res[0]); o.amp_hi = fkey_inv(res[1]);
  __syncthreads();
  // frequency: median of the wrapped first difference of the phase in degrees (:114-117)
  if (cached) block_select2([&](unsigned long long i) { return kpd[i]; }, n2, (n2 - 1) / 2, n2 / 2, h0, h1, res, sh_rank);
  else block_select2([&](unsigned long long i) {
                       return fkey(wrapped_diff(y[(a + i) * M + p.kph], y[(a + i + 1) * M + p.kph]));
                     }, n2, (n2 - 1) / 2, n2 / 2, h0, h1, res, sh_rank);
  o.pd_lo = fkey_inv(res[0]); o.pd_hi = fkey_inv(res[1]);
  o.sat = (uint32_t)sh_sat; o.pad = 0;
  if (threadIdx.x == 0) *(PulseOut*)((unsigned char*)outp + pi * STRIDE) = o;
  }
}
__global__ void __launch_bounds__(128) k_pulse_stats(const float2* __restrict__ y, long long M, double sat_level, unsigned long long row0,
                                                     const PulseIn* __restrict__ in, PulseOut* __restrict__ outp,
                                                     const unsigned long long* __restrict__ npulses_dev,
                                                     unsigned long long n_host, unsigned long long cap) {
  pulse_stats_body<(int)sizeof(PulseIn)>(y, M, sat_level, row0, in, outp, npulses_dev, n_host, cap, nullptr, 0);
}
// records laid out as {PulseIn, PulseOut} pairs (the one-GPU extractor's device-side pulse list)
__global__ void __launch_bounds__(128) k_pulse_stats_rec(const float2* __restrict__ y, long long M, double sat_level, void* recs,
                                                         const unsigned long long* __restrict__ npulses_dev, unsigned long long cap,
                                                         const uint2* __restrict__ summ, int chunk_rows) {
  pulse_stats_body<(int)(sizeof(PulseIn) + sizeof(PulseOut))>(y, M, sat_level, 0ull, (const PulseIn*)recs,
                                                             (PulseOut*)((unsigned char*)recs + sizeof(PulseIn)), npulses_dev, 0ull, cap,
                                                             summ, chunk_rows);
}

// ---- host driver ---------------------------------------------------------------------------------------
// Stages shared by the one-GPU extractor (pdw_extract) and the time-sharded one (chz_pdw_shard_*, where
// the host sums the histograms of all shards between hist and select).
static int pdw_buffers(::chz* h) {
  const int M = (int)h->M;
  CHZ_CUDA(h->pdw_hist.reserve((size_t)M * 2 * kBins * sizeof(uint32_t)));
  CHZ_CUDA(h->pdw_sel.reserve((size_t)M * sizeof(SelState)));
  CHZ_CUDA(h->pdw_thr.reserve((size_t)M * sizeof(Thr)));
  CHZ_CUDA(h->pdw_code.reserve((size_t)M));
  CHZ_CUDA(h->pdw_nf.reserve((size_t)M * sizeof(double)));
  return CHZ_OK;
}

static void threshold_scales(const chz_pdw_params_t* prm, double* scale, double* scale_lo) {
  *scale = std::pow(10.0, prm->snr_threshold_db / 10.0);
  // create_pdws.m:47: TRAILING_EDGE_THRESHOLD = NOISE_FLOOR*10^(3/10); never above the leading threshold
  const bool hyst = prm->use_trailing_threshold != 0 && prm->trailing_snr_threshold_db < prm->snr_threshold_db;
  *scale_lo = hyst ? std::pow(10.0, prm->trailing_snr_threshold_db / 10.0) : *scale;
}

// arguments of one select: the two middle ranks of total_rows values; prm != NULL on the last pass: noise floor ->
// nf_dev and thresholds -> h->pdw_thr in the same step
static int make_sel_args(::chz* h, uint64_t total_rows, const chz_pdw_params_t* prm, double* nf_dev, SelArgs* sa) {
  if (total_rows == 0 || total_rows > 0xFFFFFFFFull) return CHZ_EINVAL;
  sa->rank_lo = (uint32_t)((total_rows - 1) / 2);
  sa->rank_hi = (uint32_t)(total_rows / 2);
  sa->scale = sa->scale_lo = 0.0;
  if (prm) threshold_scales(prm, &sa->scale, &sa->scale_lo);
  sa->thr = prm ? (Thr*)h->pdw_thr.p : nullptr;
  sa->nf = nf_dev;
  return CHZ_OK;
}

// local histogram of one radix pass (:73); pass 0 clears the table first (k_select clears it after every pass)
static int pdw_hist_pass(::chz* h, const float2* y, uint64_t nrows, int pass) {
  const int M = (int)h->M;
  cudaStream_t st = h->stream;
  int rc = pdw_buffers(h);
  if (rc) return rc;
  uint32_t* d_hist = (uint32_t*)h->pdw_hist.p;
  if (pass == 0) CHZ_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)M * 2 * kBins * sizeof(uint32_t), st));
  if (nrows == 0) return CHZ_OK;
  // exactly one wave: a second, partial wave doubles the pass time on 100 ms files.  How many blocks fit an SM is
  // asked, not assumed: the first version of the kernel (32 KB of histogram per block) was sized for 7 per SM where 6
  // fit (the system reserves 1 KB per block) and ran every pass in 1.15 waves until ncu showed it (profiles/r02l_*).
  static thread_local int occ_dev[kMaxDev] = {};
  int& occ = occ_dev[h->device % kMaxDev];
  if (!occ) {
    int nb = 0;
    CHZ_CUDA(cudaFuncSetAttribute(k_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, kH16Smem));
    CHZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_hist, 256, kH16Smem));
    occ = nb > 0 ? nb : 1;
  }
  const int slabs = (M + 15) / 16;
  const int rpw = (M < 16 && 16 % M == 0) ? 16 / M : 1;   // matrix rows per 16-lane row of the block
  const long long wave = std::max<long long>(1, ((long long)h->sm_count * occ) / slabs);
  long long ychunks = wave;
  static const int hc_env = std::getenv("CHZ_PDW_HIST_CHUNKS") ? std::atoi(std::getenv("CHZ_PDW_HIST_CHUNKS")) : 0;   // tuning aid
  if (hc_env > 0) ychunks = hc_env;
  const long long max_chunks = (long long)((nrows + 255) / 256);
  if (ychunks > max_chunks) ychunks = max_chunks;
  if (ychunks < 1) ychunks = 1;
  const long long need = (long long)((nrows + 65534) / 65535);   // 16-bit counters: at most 65 535 rows per block, in whole waves
  if (ychunks < need) ychunks = (need + wave - 1) / wave * wave;
  if (ychunks > 65535) ychunks = 65535;
  if ((long long)((nrows + ychunks - 1) / ychunks) > 65535) return CHZ_EINVAL;   // > 4.29e9 rows: not reachable (uint32 ranks)
  k_hist<<<dim3(slabs, (unsigned)ychunks), 256, kH16Smem, st>>>(y, (long long)nrows, M, pass, rpw, (const SelState*)h->pdw_sel.p, d_hist);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}

// fix the next bits of both middle order statistics from the (summed) histogram; total_rows = rows of the
// WHOLE recording

// prm != NULL on the last pass: noise floor -> nf_dev and thresholds -> h->pdw_thr in the same launch
static int pdw_select_pass(::chz* h, int pass, uint64_t total_rows, const chz_pdw_params_t* prm = nullptr, double* nf_dev = nullptr) {
  SelArgs sa;
  const int rc = make_sel_args(h, total_rows, prm, nf_dev, &sa);
  if (rc) return rc;
  k_select<<<h->M, 256, 0, h->stream>>>((uint32_t*)h->pdw_hist.p, (SelState*)h->pdw_sel.p, pass, sa);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}

// noise floor and thresholds (:73-75) on the device (k_thresholds).  fetch: also bring the noise floor to
// h->noise_floor now (one synchronisation); otherwise the caller copies it with its next transfer.
static int pdw_thresholds(::chz* h, const chz_pdw_params_t* prm, bool fetch) {
  const int M = (int)h->M;
  cudaStream_t st = h->stream;
  double scale, scale_lo;
  threshold_scales(prm, &scale, &scale_lo);
  k_thresholds<<<(M + 127) / 128, 128, 0, st>>>((const SelState*)h->pdw_sel.p, M, scale, scale_lo, (Thr*)h->pdw_thr.p,
                                                (double*)h->pdw_nf.p);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  h->noise_floor.assign(M, NAN);
  if (fetch) {
    CHZ_CUDA(cudaMemcpyAsync(h->noise_floor.data(), h->pdw_nf.p, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
    CHZ_CUDA(cudaStreamSynchronize(st));
  }
  return CHZ_OK;
}

// edge events (:79-96) of y[nrows][M]; rows in the events are 1-based and offset by row_offset;
// entry_host: per natural channel state on entry (NULL = inactive).  The event counter sits at the head of
// the device event buffer so that one transfer brings the count and (normally) every event; nf_fetch: the
// noise floor rides along in front of the same synchronisation.
static int pdw_detect(::chz* h, const float2* y, uint64_t nrows, uint64_t row_offset, const uint8_t* entry_host,
                      std::vector<unsigned long long>& ev, bool nf_fetch) {
  NvtxRange nvtx_range("chz:pdw:detect(events)");
  const int M = (int)h->M;
  cudaStream_t st = h->stream;
  ev.clear();
  auto fetch_nf = [&]() -> int {   // after the launches: a copy to pageable memory may block the host
    if (nf_fetch) CHZ_CUDA(cudaMemcpyAsync(h->noise_floor.data(), h->pdw_nf.p, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
    return CHZ_OK;
  };
  if (nrows == 0) {
    if (fetch_nf()) return CHZ_ECUDA;
    CHZ_CUDA(cudaStreamSynchronize(st));
    return CHZ_OK;
  }
  uint8_t* d_entry = nullptr;
  if (entry_host) {
    d_entry = (uint8_t*)h->pdw_code.p;
    CHZ_CUDA(cudaMemcpyAsync(d_entry, entry_host, (size_t)M, cudaMemcpyHostToDevice, st));
  }
  const int chunk_rows = 64;
  const long long nchunks = ((long long)nrows + chunk_rows - 1) / chunk_rows;
  const int lanes_ch = M < 32 ? M : 32, streams = 32 / lanes_ch, ch_groups = (M + 31) / 32;
  const long long warps = ((nchunks + streams - 1) / streams) * ch_groups;
  const long long blocks = (warps + 7) / 8;
  constexpr unsigned long long kFirst = 8191;   // events fetched together with the count (64 KB)
  for (;;) {
    const unsigned long long cap = h->pdw_ev_cap;
    CHZ_CUDA(h->pdw_ev.reserve((cap + 1) * sizeof(unsigned long long)));
    unsigned long long* d_cnt = (unsigned long long*)h->pdw_ev.p;
    unsigned long long* d_ev = d_cnt + 1;
    CHZ_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st));
    k_detect<false><<<(unsigned)blocks, 256, 0, st>>>(y, (long long)nrows, M, (const Thr*)h->pdw_thr.p, chunk_rows, d_entry,
                                                      (unsigned long long)row_offset, d_ev, cap, d_cnt, -1, nullptr);
    h->launches++;
    CHZ_CUDA(cudaGetLastError());
    if (fetch_nf()) return CHZ_ECUDA;
    const unsigned long long first = cap < kFirst ? cap : kFirst;
    ev.resize(first + 1);
    CHZ_CUDA(cudaMemcpyAsync(ev.data(), d_cnt, (first + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CHZ_CUDA(cudaStreamSynchronize(st));
    const unsigned long long nev = ev[0];
    if (nev <= cap) {
      ev.erase(ev.begin());
      ev.resize(nev);
      if (nev > first) {
        CHZ_CUDA(cudaMemcpyAsync(ev.data() + first, d_ev + first, (nev - first) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CHZ_CUDA(cudaStreamSynchronize(st));
      }
      break;
    }
    h->pdw_ev_cap = nev + nev / 4;   // rerun with room for everything
  }
  return CHZ_OK;
}

// pair edges per channel in time order (sorts ev); a pulse still open at the end is dropped (:135)
// LSD radix sort of the event keys (52 significant bits: 12 of channel, 39 of row, 1 edge flag): four 13-bit
// passes.  std::sort took 1.0 ms of a 5 ms extraction on 17.7 k events.
static void sort_events(std::vector<unsigned long long>& ev) {
  if (ev.size() < 64) { std::sort(ev.begin(), ev.end()); return; }
  std::vector<unsigned long long> tmp(ev.size());
  unsigned long long* src = ev.data();
  unsigned long long* dst = tmp.data();
  unsigned long long all = 0;
  for (unsigned long long e : ev) all |= e;
  for (int shift = 0; shift < 64; shift += 13) {
    if ((all >> shift) == 0) break;
    size_t cnt[8192 + 1] = {0};
    for (size_t i = 0; i < ev.size(); i++) cnt[((src[i] >> shift) & 8191) + 1]++;
    for (int i = 0; i < 8192; i++) cnt[i + 1] += cnt[i];
    for (size_t i = 0; i < ev.size(); i++) dst[cnt[(src[i] >> shift) & 8191]++] = src[i];
    std::swap(src, dst);
  }
  if (src != ev.data()) std::copy(src, src + ev.size(), ev.data());
}

static void pdw_pair(std::vector<unsigned long long>& ev, uint32_t M, bool phase_bug, std::vector<chz_pulse_t>& pulses) {
  sort_events(ev);
  pulses.clear();
  const uint32_t kbug = (uint32_t)((0 + (M + 1) / 2) % M);   // natural channel of shifted column 1 (:114)
  for (size_t i = 0; i + 1 < ev.size(); i++) {
    const unsigned long long e0 = ev[i], e1 = ev[i + 1];
    if ((e0 & 1ull) == 0 && (e1 & 1ull) == 1 && (e0 >> 40) == (e1 >> 40)) {
      chz_pulse_t p;
      memset(&p, 0, sizeof p);
      p.toa_row = (e0 & ((1ull << 40) - 1)) >> 1;
      p.end_row = (e1 & ((1ull << 40) - 1)) >> 1;
      const uint32_t c = (uint32_t)(e0 >> 40);
      p.channel_natural = (c + (M + 1) / 2) % M;
      p.col = p.channel_natural;
      p.col_phase = phase_bug ? kbug : p.channel_natural;
      pulses.push_back(p);
      i++;
    }
  }
}

// one record (:97-128) from a pulse's edges and its device-side statistics
static chz_pdw_t make_record(const ::chz* h, const chz_pdw_params_t* prm, uint32_t k, uint64_t toa_row, uint64_t end_row,
                             const PulseOut& o) {
  const int M = (int)h->M;
  const double fs_dec = prm->fs_sps / (double)h->D;        // :62
  chz_pdw_t r;
  memset(&r, 0, sizeof r);
  const uint32_t c = (k + (uint32_t)(M / 2)) % (uint32_t)M;
  const double bin_freq = ((double)c - (double)(M / 2)) * prm->fs_sps / (double)M;   // :42 on shifted columns
  const double nf = h->noise_floor[k];
  const double med_pd = 0.5 * ((double)o.pd_lo + (double)o.pd_hi);
  r.toa_s = ((double)toa_row / fs_dec) + prm->t0;                                     // :98
  r.amp = 0.5 * ((double)o.amp_lo + (double)o.amp_hi);                                // :101
  r.snr_db = 10.0 * std::log10(r.amp / nf);                                           // :105
  r.pw_s = (double)(end_row - toa_row) / fs_dec;                                      // :110
  r.freq_hz = (prm->fc_hz + bin_freq) + (fs_dec / (360.0 / med_pd));                  // :80, :122
  r.noise_floor = nf;
  r.channel = c; r.channel_natural = k;
  r.toa_row = toa_row; r.end_row = end_row; r.saturated = o.sat;
  return r;
}

// per-pulse medians and saturation (:97-122) and the records (:97-128) for pulses whose rows all lie in
// y (leading dimension ld, row 0 = 1-based row row_offset + 1 of the run; columns p.col / p.col_phase)
static int pdw_records(::chz* h, const chz_pdw_params_t* prm, const float2* y, uint64_t ld, uint64_t row_offset,
                       const chz_pulse_t* pulses, size_t n, chz_pdw_t* out) {
  if (n == 0) return CHZ_OK;
  if (h->noise_floor.size() != h->M) return CHZ_ESTATE;
  NvtxRange nvtx_range("chz:pdw:stats");
  cudaStream_t st = h->stream;
  std::vector<PulseIn> pin(n);
  for (size_t i = 0; i < n; i++) {
    pin[i].toa = pulses[i].toa_row; pin[i].end = pulses[i].end_row;
    pin[i].k = pulses[i].col; pin[i].kph = pulses[i].col_phase;
    if (pin[i].toa <= row_offset || pin[i].end < pin[i].toa || pin[i].k >= ld || pin[i].kph >= ld) return CHZ_EINVAL;
  }
  CHZ_CUDA(h->pdw_pin.reserve(n * sizeof(PulseIn)));
  CHZ_CUDA(h->pdw_pout.reserve(n * sizeof(PulseOut)));
  PulseIn* d_pin = (PulseIn*)h->pdw_pin.p;
  PulseOut* d_pout = (PulseOut*)h->pdw_pout.p;
  CHZ_CUDA(cudaMemcpyAsync(d_pin, pin.data(), n * sizeof(PulseIn), cudaMemcpyHostToDevice, st));
  k_pulse_stats<<<(unsigned)n, 128, 0, st>>>(y, (long long)ld, prm->sat_level, (unsigned long long)row_offset, d_pin, d_pout,
                                             nullptr, (unsigned long long)n, (unsigned long long)n);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  std::vector<PulseOut> pout(n);
  CHZ_CUDA(cudaMemcpyAsync(pout.data(), d_pout, n * sizeof(PulseOut), cudaMemcpyDeviceToHost, st));
  CHZ_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < n; i++) out[i] = make_record(h, prm, pulses[i].channel_natural, pulses[i].toa_row, pulses[i].end_row, pout[i]);
  return CHZ_OK;
}

// One-GPU extractor without a host round trip between its stages: median passes -> thresholds (fused into the last
// select) -> k_detect<true> (finished pulses, device-side list and count) -> k_pulse_stats striding over that list ->
// ONE copy of count + noise floor + pulses + statistics into pinned memory -> ONE synchronisation.  The host then
// only builds the records and orders them as the script does (:79,85: shifted channel ascending, then time).
// Returns 1 when the kernel asked for the event path (a sample exactly on a single representable threshold).
struct PulseRec { PulseIn in; PulseOut out; };
constexpr unsigned long long kStageFirst = 2048;     // pulse records fetched together with the count

static int pdw_extract_fast(::chz* h, const chz_pdw_params_t* prm, const float2* y, uint64_t nrows) {
  const int M = (int)h->M;
  cudaStream_t st = h->stream;
  int rc = pdw_buffers(h);
  if (rc) return rc;
  const size_t head = 16 + (size_t)M * sizeof(double);                    // [count, fallback flag][noise floor]
  if (!h->pdw_stage_host || h->pdw_stage_bytes < head + kStageFirst * sizeof(PulseRec)) {
    if (h->pdw_stage_host) cudaFreeHost(h->pdw_stage_host);
    h->pdw_stage_host = nullptr;
    h->pdw_stage_bytes = head + kStageFirst * sizeof(PulseRec);
    CHZ_CUDA(cudaMallocHost(&h->pdw_stage_host, h->pdw_stage_bytes));
  }
  static const bool gtrace = std::getenv("CHZ_PDW_TRACE") != nullptr;   // GPU time of each stage (CUDA events), debug aid
  cudaEvent_t tev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool capturing = false;
  auto mark = [&](int i) { if (gtrace && !capturing) { if (!tev[i]) cudaEventCreate(&tev[i]); cudaEventRecord(tev[i], st); } };
  static const int chunk_env = std::getenv("CHZ_PDW_CHUNK_ROWS") ? std::atoi(std::getenv("CHZ_PDW_CHUNK_ROWS")) : 0;   // tuning aid (multiple of 16)
  const int chunk_rows = chunk_env > 0 ? chunk_env : 64;
  const long long nchunks = ((long long)nrows + chunk_rows - 1) / chunk_rows;
  const int lanes_ch = M < 32 ? M : 32, streams = 32 / lanes_ch, ch_groups = (M + 31) / 32;
  const long long warps = ((nchunks + streams - 1) / streams) * ch_groups;
  const long long blocks = (warps + 7) / 8;
  CHZ_CUDA(h->pdw_ev.reserve((size_t)nchunks * M * sizeof(uint2)));       // chunk summaries (the event list is not used here)
  uint2* d_summ = (uint2*)h->pdw_ev.p;
  NvtxRange nvtx_range("chz:pdw:median+detect+stats");
  const int kbug = prm->reproduce_phase_bug ? (int)((0 + (M + 1) / 2) % M) : -1;   // natural channel of shifted column 1 (:114)
  bool first_try = true;
  for (;;) {
    const unsigned long long cap = h->pdw_pulse_cap;
    CHZ_CUDA(h->pdw_fast.reserve(head + cap * sizeof(PulseRec)));
    unsigned char* base = (unsigned char*)h->pdw_fast.p;
    unsigned long long* d_cnt = (unsigned long long*)base;
    double* d_nf = (double*)(base + 16);
    PulseRec* d_rec = (PulseRec*)(base + head);
    const unsigned long long first = std::min<unsigned long long>(cap, kStageFirst);
    // every stream operation of one attempt; the same sequence is either issued directly or captured into a graph
    auto enqueue = [&](bool with_median) -> int {
      int r;
      mark(0);
      if (with_median) {
        for (int pass = 0; pass < 3; pass++) {
          if ((r = pdw_hist_pass(h, y, nrows, pass))) return r;
          if (pass < 2 && (r = pdw_select_pass(h, pass, nrows))) return r;
          if (pass == 0) mark(1);
        }
        // last select + thresholds, writing the noise floor next to the counters
        if ((r = pdw_select_pass(h, 2, nrows, prm, d_nf))) return r;
      } else {            // rerun with a larger list: the noise floor moved with the buffer
        CHZ_CUDA(cudaMemcpyAsync(d_nf, h->noise_floor.data(), sizeof(double) * M, cudaMemcpyHostToDevice, st));
      }
      mark(2);
      CHZ_CUDA(cudaMemsetAsync(d_cnt, 0, 16, st));
      k_detect<true><<<(unsigned)blocks, 256, 0, st>>>(y, (long long)nrows, M, (const Thr*)h->pdw_thr.p, chunk_rows, nullptr, 0ull,
                                                       (unsigned long long*)d_rec, cap, d_cnt, kbug, d_summ);
      h->launches++;
      CHZ_CUDA(cudaGetLastError());
      // PulseRec interleaves input and output, so detect writes .in of slot i and the statistics kernel .out
      static_assert(sizeof(PulseRec) == sizeof(PulseIn) + sizeof(PulseOut), "packed");
      mark(3);
      const unsigned sblocks = (unsigned)std::min<unsigned long long>(cap, (unsigned long long)h->sm_count * 8);
      k_pulse_stats_rec<<<sblocks, 128, 0, st>>>(y, (long long)M, prm->sat_level, d_rec, d_cnt, cap, d_summ, chunk_rows);
      h->launches++;
      CHZ_CUDA(cudaGetLastError());
      mark(4);
      CHZ_CUDA(cudaMemcpyAsync(h->pdw_stage_host, base, head + first * sizeof(PulseRec), cudaMemcpyDeviceToHost, st));
      mark(5);
      return CHZ_OK;
    };
    bool done = false;
    if (first_try && h->pdw_use_graph && !gtrace) {
      // everything the captured operations depend on: a change of any of it means a new capture
      unsigned char key[sizeof(h->pdw_graph_key)] = {0};
      size_t ko = 0;
      auto put = [&](const void* p, size_t nb) { memcpy(key + ko, p, nb); ko += nb; };
      const void* ptrs[] = {y, base, h->pdw_stage_host, h->pdw_hist.p, h->pdw_sel.p, h->pdw_thr.p, h->pdw_ev.p, (const void*)st};
      put(ptrs, sizeof(ptrs)); put(&nrows, sizeof(nrows)); put(&cap, sizeof(cap)); put(&chunk_rows, sizeof(chunk_rows));
      static_assert(sizeof(ptrs) + 2 * 8 + 4 + sizeof(chz_pdw_params_t) <= sizeof(key), "graph key");
      put(prm, sizeof(*prm));
      if (!h->pdw_graph || memcmp(key, h->pdw_graph_key, sizeof(key)) != 0) {
        if (h->pdw_graph) { cudaGraphExecDestroy(h->pdw_graph); h->pdw_graph = nullptr; }
        const uint64_t l0 = h->launches;
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          capturing = true;
          const int r = enqueue(true);
          capturing = false;
          const cudaError_t ce = cudaStreamEndCapture(st, &g);
          if (r == CHZ_OK && ce == cudaSuccess && g && cudaGraphInstantiate(&h->pdw_graph, g, 0) == cudaSuccess) {
            memcpy(h->pdw_graph_key, key, sizeof(key));
            h->pdw_graph_kernels = (int)(h->launches - l0);
          } else {
            h->pdw_graph = nullptr;
          }
          if (g) cudaGraphDestroy(g);
        }
        h->launches = l0;
        cudaGetLastError();
      }
      if (h->pdw_graph) {
        CHZ_CUDA(cudaGraphLaunch(h->pdw_graph, st));
        h->launches += h->pdw_graph_kernels;
        done = true;
      }
    }
    if (!done && (rc = enqueue(first_try))) return rc;
    first_try = false;
    CHZ_CUDA(cudaStreamSynchronize(st));
    if (gtrace && tev[0] && tev[1] && tev[2] && tev[3] && tev[4] && tev[5]) {
      float ms[5];
      for (int i = 0; i < 5; i++) cudaEventElapsedTime(&ms[i], tev[i], tev[i + 1]);
      std::fprintf(stderr, "[chz pdw gpu] hist0+sel0 %.1f us | hist1..sel2 %.1f | detect %.1f | stats %.1f | copy %.1f\n", ms[0] * 1e3,
                   ms[1] * 1e3, ms[2] * 1e3, ms[3] * 1e3, ms[4] * 1e3);
      for (int i = 0; i < 6; i++) { cudaEventDestroy(tev[i]); tev[i] = nullptr; }
    }
    const auto t_post = std::chrono::steady_clock::now();
    const unsigned long long* hc = (const unsigned long long*)h->pdw_stage_host;
    const unsigned long long n = hc[0];
    memcpy(h->noise_floor.data(), (const unsigned char*)h->pdw_stage_host + 16, sizeof(double) * M);
    if (hc[1]) return 1;                               // equality on a representable threshold: event path
    if (n > cap) { h->pdw_pulse_cap = n + n / 4; continue; }
    // the records usually all sit in the pinned landing zone already; only a longer list needs a second copy
    const unsigned long long got = std::min(n, first);
    const PulseRec* recs = (const PulseRec*)((const unsigned char*)h->pdw_stage_host + head);
    std::vector<PulseRec> more;
    if (n > got) {
      more.resize(n);
      memcpy(more.data(), recs, got * sizeof(PulseRec));
      CHZ_CUDA(cudaMemcpyAsync(more.data() + got, d_rec + got, (n - got) * sizeof(PulseRec), cudaMemcpyDeviceToHost, st));
      CHZ_CUDA(cudaStreamSynchronize(st));
      recs = more.data();
    }
    // The script's order (shifted channel, then time of arrival).  The device list is in no particular order across
    // channels and nearly in time order within one: a counting sort by channel, then a small sort per channel, on
    // indices -- each 100-byte record is built once, in its final place (sorting the records themselves cost more host
    // time than the detector costs GPU time).
    std::vector<uint32_t> start((size_t)M + 1, 0u), idx(n);
    for (unsigned long long i = 0; i < n; i++) start[(recs[i].in.k + (uint32_t)(M / 2)) % (uint32_t)M + 1]++;
    for (int c = 0; c < M; c++) start[c + 1] += start[c];
    {
      std::vector<uint32_t> fill(start.begin(), start.end() - 1);
      for (unsigned long long i = 0; i < n; i++) idx[fill[(recs[i].in.k + (uint32_t)(M / 2)) % (uint32_t)M]++] = (uint32_t)i;
    }
    for (int c = 0; c < M; c++)
      if (start[c + 1] - start[c] > 1)
        std::sort(idx.begin() + start[c], idx.begin() + start[c + 1], [&](uint32_t x, uint32_t y2) { return recs[x].in.toa < recs[y2].in.toa; });
    h->pdws.resize(n);
    for (unsigned long long i = 0; i < n; i++) {
      const PulseRec& q = recs[idx[i]];
      h->pdws[i] = make_record(h, prm, q.in.k, q.in.toa, q.in.end, q.out);
    }
    if (gtrace)
      std::fprintf(stderr, "[chz pdw host] %llu records built and ordered in %.1f us\n", n,
                   std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_post).count());
    return CHZ_OK;
  }
}

int pdw_extract(::chz* h, const chz_pdw_params_t* prm, const float2* y, uint64_t nrows) {
  NvtxRange nvtx_range("chz:pdws");
  const int M = (int)h->M;
  // CHZ_PDW_TRACE=1: host wall-clock of each stage on stderr (debug aid)
  static const bool trace = std::getenv("CHZ_PDW_TRACE") != nullptr;
  const bool events_only = h->pdw_event_path;   // CHZ_OPT_PDW_EVENT_PATH: always take the event path (A/B, tests)
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[chz pdw] %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  h->pdws.clear();
  h->noise_floor.assign(M, NAN);
  if (nrows == 0) return CHZ_OK;
  int rc;
  if (!events_only) {
    rc = pdw_extract_fast(h, prm, y, nrows);
    lap("fast path");
    if (rc != 1) { if (rc) h->pdws.clear(); return rc; }
    h->pdws.clear();                                 // fall through: thresholds and noise floor are already in place
  } else {
    // 1. exact per-channel median of |y| (:73)
    for (int pass = 0; pass < 3; pass++) {
      if ((rc = pdw_hist_pass(h, y, nrows, pass))) return rc;
      if ((rc = pdw_select_pass(h, pass, nrows))) return rc;
    }
    lap("median");
    // 2. thresholds (:74-75)
    if ((rc = pdw_thresholds(h, prm, false))) return rc;
  }
  // 3. edge events (:79-96); the noise floor comes back with the event count
  std::vector<unsigned long long> ev;
  if ((rc = pdw_detect(h, y, nrows, 0, nullptr, ev, events_only))) return rc;
  lap("detect");
  // 4. pulses in the reference's order: shifted channel ascending, then time
  std::vector<chz_pulse_t> pulses;
  pdw_pair(ev, (uint32_t)M, prm->reproduce_phase_bug != 0, pulses);
  lap("pair");
  if (pulses.empty()) return CHZ_OK;
  // 5./6. per-pulse statistics and records
  h->pdws.resize(pulses.size());
  rc = pdw_records(h, prm, y, (uint64_t)M, 0, pulses.data(), pulses.size(), h->pdws.data());
  if (rc) h->pdws.clear();
  lap("stats");
  return rc;
}

}  // namespace chzi

using namespace chzi;

// ---- time-sharded PDW extraction: the stages above, one shard per GPU --------------------
extern "C" {

int chz_pdw_shard_hist_dev(chz_t* h, const chz_cf32* y_dev, uint64_t nrows, int pass, uint32_t** hist_dev, uint64_t* hist_words) {
  if (!h || (!y_dev && nrows) || pass < 0 || pass > 2) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  const int rc = pdw_hist_pass(h, (const float2*)y_dev, nrows, pass);
  if (rc) return rc;
  CHZ_CUDA(cudaStreamSynchronize(h->stream));   // the caller sums the table on its own stream / communicator
  if (hist_dev) *hist_dev = (uint32_t*)h->pdw_hist.p;
  if (hist_words) *hist_words = (uint64_t)h->M * 2 * kBins;
  return CHZ_OK;
}

int chz_pdw_shard_select(chz_t* h, int pass, uint64_t total_rows) {
  if (!h || pass < 0 || pass > 2 || !h->pdw_hist.p) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  const int rc = pdw_select_pass(h, pass, total_rows);
  if (rc) return rc;
  CHZ_CUDA(cudaStreamSynchronize(h->stream));
  return CHZ_OK;
}

int chz_pdw_shard_thresholds(chz_t* h, const chz_pdw_params_t* params) {
  if (!h || !params || !h->pdw_sel.p) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  h->pdws.clear();
  return pdw_thresholds(h, params, true);
}

int chz_pdw_shard_set_noise_floor(chz_t* h, const chz_pdw_params_t* params, const double* noise_floor) {
  if (!h || !params || !noise_floor) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  const int rc = pdw_buffers(h);
  if (rc) return rc;
  const uint32_t M = h->M;
  double scale, scale_lo;
  threshold_scales(params, &scale, &scale_lo);
  std::vector<Thr> thr(M);
  for (uint32_t k = 0; k < M; k++) {     // the same bracketing as thresholds_of() on the device
    const double tl = noise_floor[k] * scale, tt = noise_floor[k] * scale_lo;
    float ge = (float)tl, le = (float)tt;
    if ((double)ge < tl) ge = std::nextafterf(ge, INFINITY);
    if ((double)le > tt) le = std::nextafterf(le, -INFINITY);
    thr[k].ge = ge; thr[k].le = le;
    squared_bounds(&thr[k]);
  }
  h->noise_floor.assign(noise_floor, noise_floor + M);
  h->pdws.clear();
  CHZ_CUDA(cudaMemcpyAsync(h->pdw_thr.p, thr.data(), sizeof(Thr) * M, cudaMemcpyHostToDevice, h->stream));
  CHZ_CUDA(cudaMemcpyAsync(h->pdw_nf.p, noise_floor, sizeof(double) * M, cudaMemcpyHostToDevice, h->stream));
  CHZ_CUDA(cudaStreamSynchronize(h->stream));
  return CHZ_OK;
}

int chz_pdw_shard_exit_state_dev(chz_t* h, const chz_cf32* y_dev, uint64_t nrows, uint8_t* code) {
  if (!h || !code || (!y_dev && nrows)) return CHZ_EINVAL;
  if (h->noise_floor.size() != h->M) return CHZ_ESTATE;
  CHZ_CUDA(cudaSetDevice(h->device));
  k_exit_state<<<(h->M + 127) / 128, 128, 0, h->stream>>>((const float2*)y_dev, (long long)nrows, (int)h->M,
                                                           (const Thr*)h->pdw_thr.p, (uint8_t*)h->pdw_code.p);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  CHZ_CUDA(cudaMemcpyAsync(code, h->pdw_code.p, h->M, cudaMemcpyDeviceToHost, h->stream));
  CHZ_CUDA(cudaStreamSynchronize(h->stream));
  return CHZ_OK;
}

int chz_pdw_shard_detect_dev(chz_t* h, const chz_cf32* y_dev, uint64_t nrows, uint64_t row_offset, const uint8_t* entry,
                             uint64_t* events, uint64_t cap, uint64_t* n) {
  if (!h || !n || (!y_dev && nrows)) return CHZ_EINVAL;
  if (h->noise_floor.size() != h->M) return CHZ_ESTATE;
  if (row_offset + nrows >= (1ull << 39)) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  std::vector<unsigned long long> ev;
  const int rc = pdw_detect(h, (const float2*)y_dev, nrows, row_offset, entry, ev, false);
  if (rc) return rc;
  *n = ev.size();
  if (ev.size() > cap) return CHZ_ECAPACITY;
  if (events && !ev.empty()) memcpy(events, ev.data(), ev.size() * sizeof(uint64_t));
  return CHZ_OK;
}

int chz_pdw_pair_events(uint64_t* events, uint64_t n, uint32_t M, uint32_t reproduce_phase_bug, chz_pulse_t* pulses,
                        uint64_t cap, uint64_t* npulses) {
  if ((!events && n) || !npulses || M == 0) return CHZ_EINVAL;
  std::vector<unsigned long long> ev(events, events + n);
  std::vector<chz_pulse_t> p;
  pdw_pair(ev, M, reproduce_phase_bug != 0, p);
  *npulses = p.size();
  if (p.size() > cap) return CHZ_ECAPACITY;
  if (pulses && !p.empty()) memcpy(pulses, p.data(), p.size() * sizeof(chz_pulse_t));
  return CHZ_OK;
}

int chz_pdw_shard_records_dev(chz_t* h, const chz_pdw_params_t* params, const chz_cf32* y_dev, uint64_t ld,
                              uint64_t row_offset, const chz_pulse_t* pulses, uint64_t n, chz_pdw_t* out) {
  if (!h || !params || ((!y_dev || !pulses || !out) && n) || ld == 0) return CHZ_EINVAL;
  CHZ_CUDA(cudaSetDevice(h->device));
  return pdw_records(h, params, (const float2*)y_dev, ld, row_offset, pulses, (size_t)n, out);
}

}  // extern "C"
