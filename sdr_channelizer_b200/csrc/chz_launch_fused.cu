// Launchers of the fused K1+K2+K3 kernel (k_chan_fused) for every (M, P) instantiation.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "chz_internal.h"
#include "chz_kernels.cuh"
#include "chz_launch.h"

namespace chzi {

template <int M, int P, bool IN16>
static int launch_fused(::chz* h, ChanParams prm, cudaStream_t st) {
  typedef FusedCfg<M, P> CF;
  auto kern = k_chan_fused<M, P, IN16>;
  static thread_local int blocks_per_sm_dev[kMaxDev] = {0};   // launch geometry is cached per device
  int& blocks_per_sm = blocks_per_sm_dev[h->device % kMaxDev];
  if (!blocks_per_sm) {
    CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM));
    int nb = 0;
    CHZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, CF::NT, CF::SMEM));
    blocks_per_sm = nb > 0 ? nb : 1;
  }
  const LaunchPlan lp = plan_spans(h, prm.nrows, P, CF::G, blocks_per_sm);
  prm.span_rows = lp.span_rows;
  prm.spans_per_phase = lp.spans_per_phase;
  kern<<<lp.grid, CF::NT, CF::SMEM, st>>>(prm);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}

#define CHZ_FUSED_P(MV, IN16V)                                                   \
  switch (h->P) {                                                                \
    case 8: return launch_fused<MV, 8, IN16V>(h, prm, st);                       \
    case 12: return launch_fused<MV, 12, IN16V>(h, prm, st);                     \
    case 16: return launch_fused<MV, 16, IN16V>(h, prm, st);                     \
    default: return 1;                                                           \
  }

// returns 1 when no fused instantiation exists for (M, P)
template <bool IN16>
static int launch_fused_dispatch(::chz* h, const ChanParams& prm, cudaStream_t st) {
  switch (h->M) {
    case 8: CHZ_FUSED_P(8, IN16)
    case 16: CHZ_FUSED_P(16, IN16)
    case 32: CHZ_FUSED_P(32, IN16)
    case 64: CHZ_FUSED_P(64, IN16)
    case 128: CHZ_FUSED_P(128, IN16)
    case 256: CHZ_FUSED_P(256, IN16)
    case 512: CHZ_FUSED_P(512, IN16)
    case 56: CHZ_FUSED_P(56, IN16)
    case 560: CHZ_FUSED_P(560, IN16)
    default: return 1;
  }
}


int launch_fused_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st) {
  return in16 ? launch_fused_dispatch<true>(h, prm, st) : launch_fused_dispatch<false>(h, prm, st);
}

}  // namespace chzi
