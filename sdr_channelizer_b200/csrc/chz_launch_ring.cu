// Launcher of the large-M fused ring kernel (k_chan_ring_ws, chz_ring.cuh): M = 1024, P in {8, 12, 16}.
#include <algorithm>
#include <cstdlib>

#include "chz_internal.h"
#include "chz_ring.cuh"
#ifdef CHZ_EXPERIMENTS
#include "chz_ring_exp.cuh"
#endif
#include "chz_launch.h"

namespace chzi {

template <typename K>
static int launch_one(::chz* h, K kern, bool& attr, long long grid, int threads, int smem, const ChanParams& prm,
                      const ring::RingParams& rp, cudaStream_t st) {
  if (!attr) {
    CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  kern<<<(unsigned)grid, threads, smem, st>>>(prm, rp);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}

template <int P, bool IN16, int UNPACK>
static int launch_ring(::chz* h, const ChanParams& prm, cudaStream_t st) {
  typedef ring::Smem<IN16> SM;
  ring::RingParams rp;
  const long long os = prm.os;
  rp.a_lo = prm.row_base / os;
  const long long a_hi = (prm.row_base + prm.nrows - 1) / os + 1;
  rp.nsteps = (a_hi - rp.a_lo + ring::kR - 1) / ring::kR;
  rp.twn = h->d_twn;
  rp.dbg = h->ring_dbg;
  // one persistent CTA per SM; a CTA's run starts with a 16-frame warm-up, so short calls use fewer CTAs
  const long long grid = std::min<long long>(h->sm_count, std::max<long long>(1, rp.nsteps / h->ring_min_steps));
  static thread_local bool attr_dev[4][kMaxDev] = {};
#ifdef CHZ_EXPERIMENTS
  if (h->ring_variant == 1)      // every warp filters, then transforms
    return launch_one(h, ring::k_chan_ring<P, IN16, UNPACK>, attr_dev[1][h->device % kMaxDev], grid, ring::kNT, SM::TOTAL, prm, rp, st);
  if (h->ring_variant == 3)      // 16 FFT warps (two rows per group of four) next to the 8 FIR warps: 267 against 310 GS/s
    return launch_one(h, ring::k_chan_ring_ws<P, IN16, UNPACK, 16>, attr_dev[3][h->device % kMaxDev], grid, 768, SM::TOTAL, prm, rp, st);
  if (h->ring_variant == 2)      // 1024 threads, one branch each
    return launch_one(h, ring::k_chan_ring1k<P, IN16, UNPACK>, attr_dev[2][h->device % kMaxDev], grid, ring::kNT1k, SM::TOTAL, prm, rp, st);
#endif
  return launch_one(h, ring::k_chan_ring_ws<P, IN16, UNPACK>, attr_dev[0][h->device % kMaxDev], grid, ring::kNT, SM::TOTAL, prm, rp, st);
}

template <bool IN16, int UNPACK>
static int launch_ring_p(::chz* h, const ChanParams& prm, cudaStream_t st) {
  switch (h->P) {
    case 8: return launch_ring<8, IN16, UNPACK>(h, prm, st);
    case 12: return launch_ring<12, IN16, UNPACK>(h, prm, st);
    case 16: return launch_ring<16, IN16, UNPACK>(h, prm, st);
    default: return 1;
  }
}

bool ring_available(const ::chz* h) { return !h->generic && h->M == 1024 && (h->P == 8 || h->P == 12 || h->P == 16); }

int launch_ring_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st) {
  if (!ring_available(h)) return 1;
  // unpack policy 1 (I2F.S16 for I, shift + I2FP for Q) measured 4 % faster than 0 (two I2F.S16, all on the conversion
  // pipe behind the MIO queue); 0 stays selectable (CHZ_RING_UNPACK=0) for A/B runs
  if (h->ring_unpack == 0) return in16 ? launch_ring_p<true, 0>(h, prm, st) : launch_ring_p<false, 0>(h, prm, st);
  return in16 ? launch_ring_p<true, 1>(h, prm, st) : launch_ring_p<false, 1>(h, prm, st);
}

}  // namespace chzi
