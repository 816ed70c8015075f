// Launchers of the large-M / warp-specialised experiment kernels (CHZ_OPT_FORCE_PATH 3..10).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "chz_internal.h"
#include "chz_launch.h"

#ifdef CHZ_EXPERIMENTS
#include "chz_kernels_exp.cuh"

namespace chzi {

// M = 1024 on CTA pairs (decimation-in-time split over DSMEM).
template <int P, bool IN16>
static int launch_dit2(::chz* h, ChanParams prm, cudaStream_t st) {
  typedef Dit2Cfg<P> DC;
  auto kern = k_chan_dit2<P, IN16>;
  static thread_local bool attr_dev[kMaxDev] = {false};
  bool& attr = attr_dev[h->device % kMaxDev];
  if (!attr) {
    CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DC::SMEM));
    attr = true;
  }
  const int nclusters = h->sm_count / 2;                       // one 512-thread CTA per SM
  const LaunchPlan lp = plan_spans(h, prm.nrows, P, 1, 1, nclusters);
  prm.span_rows = lp.span_rows;
  prm.spans_per_phase = lp.spans_per_phase;
  kern<<<lp.grid.x * 2, 512, DC::SMEM, st>>>(prm);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}
template <bool IN16>
static int launch_dit2_dispatch(::chz* h, const ChanParams& prm, cudaStream_t st) {
  switch (h->P) {
    case 8: return launch_dit2<8, IN16>(h, prm, st);
    case 12: return launch_dit2<12, IN16>(h, prm, st);
    case 16: return launch_dit2<16, IN16>(h, prm, st);
    default: return 1;
  }
}
bool dit2_available(const ::chz* h) { return h->M == 1024 && (h->P == 8 || h->P == 12 || h->P == 16); }

// Warp-specialised fused kernel (M = 64).
template <int P, bool IN16>
static int launch_ws(::chz* h, ChanParams prm, cudaStream_t st) {
  typedef WsCfg<64, P> WC;
  auto kern = k_chan_ws<64, P, IN16>;
  static thread_local int blocks_per_sm_dev[kMaxDev] = {0};   // launch geometry is cached per device
  int& blocks_per_sm = blocks_per_sm_dev[h->device % kMaxDev];
  if (!blocks_per_sm) {
    CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WC::SMEM));
    int nb = 0;
    CHZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, WC::SMEM));
    blocks_per_sm = nb > 0 ? nb : 1;
  }
  const LaunchPlan lp = plan_spans(h, prm.nrows, P, 2, blocks_per_sm);
  prm.span_rows = lp.span_rows;
  prm.spans_per_phase = lp.spans_per_phase;
  kern<<<lp.grid, 256, WC::SMEM, st>>>(prm);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}
template <bool IN16>
static int launch_ws_dispatch(::chz* h, const ChanParams& prm, cudaStream_t st) {
  switch (h->P) {
    case 8: return launch_ws<8, IN16>(h, prm, st);
    case 12: return launch_ws<12, IN16>(h, prm, st);
    case 16: return launch_ws<16, IN16>(h, prm, st);
    default: return 1;
  }
}
bool ws_available(const ::chz* h) { return h->M == 64 && (h->P == 8 || h->P == 12 || h->P == 16); }

// Cluster path (M = 1024, 2048, 4096): returns 1 when there is no instantiation for (M, P).
template <int M, int P, bool IN16, int TPC, bool PIPE>
static int launch_cluster(::chz* h, ChanParams prm, cudaStream_t st) {
  typedef ClusterCfg<M, P, TPC> CC;
  if constexpr (!CC::ok) {
    return 1;
  } else {
    auto kern = k_chan_cluster<M, P, IN16, TPC, PIPE>;
    constexpr int NSLOT = PIPE ? 4 : 2;
    static thread_local int nclusters_dev[kMaxDev] = {0};
    int& nclusters = nclusters_dev[h->device % kMaxDev];
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CC::C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(TPC); cfg.dynamicSmemBytes = CC::SMEM; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
    if (!nclusters) {
      CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CC::SMEM));
      if (CC::C > 8) CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cfg.gridDim = dim3((unsigned)(h->sm_count * (TPC == 512 ? 1 : 2) / CC::C * CC::C));
      int n = 0;
      CHZ_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      nclusters = n > 0 ? n : 1;
      if (std::getenv("CHZ_TRACE_LAUNCH")) std::fprintf(stderr, "[chz] cluster kernel M=%d TPC=%d: %d clusters of %d CTAs resident\n", M, TPC, nclusters, CC::C);
    }
    const LaunchPlan lp = plan_spans(h, prm.nrows, P, 1, 1, nclusters);
    prm.span_rows = lp.span_rows;
    prm.spans_per_phase = lp.spans_per_phase;
    const unsigned ncl = lp.grid.x;   // <= nclusters
    CHZ_CUDA(h->cluster_ring.reserve((size_t)nclusters * NSLOT * P * M * sizeof(float2)));
    cfg.gridDim = dim3(ncl * CC::C);
    float2* ring = (float2*)h->cluster_ring.p;
    CHZ_CUDA(cudaLaunchKernelEx(&cfg, kern, prm, ring));
    h->launches++;
    return CHZ_OK;
  }
}

template <bool IN16, int TPC, bool PIPE>
static int launch_cluster_dispatch(::chz* h, const ChanParams& prm, cudaStream_t st) {
#define CHZ_CL_P(MV)                                                        \
  switch (h->P) {                                                           \
    case 8: return launch_cluster<MV, 8, IN16, TPC, PIPE>(h, prm, st);      \
    case 16: return launch_cluster<MV, 16, IN16, TPC, PIPE>(h, prm, st);    \
    default: return 1;                                                      \
  }
  switch (h->M) {
    case 1024: CHZ_CL_P(1024)
    case 2048: CHZ_CL_P(2048)
    case 4096: CHZ_CL_P(4096)
    default: return 1;
  }
#undef CHZ_CL_P
}

bool cluster_available(const ::chz* h, int tpc) {
  if (h->M != 1024 && h->M != 2048 && h->M != 4096) return false;
  const uint32_t C = h->M / tpc;
  if (tpc == 256) return h->P == 16 && C <= 16;
  return (h->P == 8 || h->P == 16) && h->P % C == 0;
}

// DSMEM cluster kernel (st.async + mbarrier hand-off): returns 1 when there is no instantiation for (M, P).
template <int M, int P, bool IN16>
static int launch_dsm(::chz* h, ChanParams prm, cudaStream_t st) {
  typedef DsmCfg<M, P> DC;
  if constexpr (!DC::ok) {
    return 1;
  } else {
    auto kern = k_chan_dsm<M, P, IN16>;
    static thread_local int nclusters_dev[kMaxDev] = {0};
    int& nclusters = nclusters_dev[h->device % kMaxDev];
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = DC::C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(DC::TPC); cfg.dynamicSmemBytes = DC::SMEM; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
    if (!nclusters) {
      CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DC::SMEM));
      if (DC::C > 8) CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cfg.gridDim = dim3((unsigned)(h->sm_count * 2 / DC::C * DC::C));
      int n = 0;
      CHZ_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      nclusters = n > 0 ? n : 1;
      if (std::getenv("CHZ_TRACE_LAUNCH")) std::fprintf(stderr, "[chz] dsm kernel M=%d: %d clusters of %d CTAs resident\n", M, nclusters, DC::C);
    }
    const LaunchPlan lp = plan_spans(h, prm.nrows, P, 1, 1, nclusters);
    prm.span_rows = lp.span_rows;
    prm.spans_per_phase = lp.spans_per_phase;
    cfg.gridDim = dim3(lp.grid.x * DC::C);   // lp.grid.x <= nclusters
    CHZ_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
    h->launches++;
    return CHZ_OK;
  }
}
template <bool IN16>
static int launch_dsm_dispatch(::chz* h, const ChanParams& prm, cudaStream_t st) {
  if (h->P != 16) return 1;
  switch (h->M) {
    case 1024: return launch_dsm<1024, 16, IN16>(h, prm, st);
    case 2048: return launch_dsm<2048, 16, IN16>(h, prm, st);
    case 4096: return launch_dsm<4096, 16, IN16>(h, prm, st);
    default: return 1;
  }
}
bool dsm_available(const ::chz* h) { return (h->M == 1024 || h->M == 2048 || h->M == 4096) && h->P == 16; }

// Pipelined split path (M = 1024, 2048, 4096): one persistent launch, FIR and in-place FFT tasks from one
// ordered ticket queue (k_chan_pipe).  Returns 1 when there is no instantiation for (M, P).
template <int M, int P, bool IN16>
static int launch_pipe(::chz* h, ChanParams prm, cudaStream_t st) {
  auto kern = k_chan_pipe<M, P, IN16>;
  constexpr int ROWS = 4096 / M;
  const size_t smem = (size_t)(2 * ROWS * RowStride<M>::value) * sizeof(float2);
  static thread_local int blocks_per_sm_dev[kMaxDev] = {0};   // launch geometry is cached per device
  int& blocks_per_sm = blocks_per_sm_dev[h->device % kMaxDev];
  if (!blocks_per_sm) {
    CHZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CHZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, smem));
    blocks_per_sm = nb > 0 ? nb : 1;
  }
  long long max_blocks = (long long)h->sm_count * blocks_per_sm;
  if (h->pipe_blocks > 0 && h->pipe_blocks < max_blocks) max_blocks = h->pipe_blocks;
  // row groups: os * span_rows rows, span_rows a multiple of P near h->pipe_span_rows
  long long sr = ((long long)h->pipe_span_rows + P - 1) / P * P;
  if (sr < 2 * P) sr = 2 * P;
  const long long rows_per_phase = (prm.nrows + prm.os - 1) / prm.os + 1;   // +1: a phase may start one row early
  prm.span_rows = (int)sr;
  prm.spans_per_phase = (rows_per_phase + sr - 1) / sr;
  PipeParams pp;
  memset(&pp, 0, sizeof pp);
  const long long group_rows = (long long)prm.os * sr;
  pp.ngroups_fir = (int)prm.spans_per_phase;
  pp.ngroups_fft = (int)((prm.nrows + group_rows - 1) / group_rows);
  pp.tpg = prm.os * (M / 256);
  long long sub = 32768 / M;                      // ~32 Ki samples per FFT task, like a FIR task
  if (sub < ROWS) sub = ROWS;
  pp.sub_rows = (int)sub;
  pp.tsub = (int)((group_rows + sub - 1) / sub);
  const int slot_len = pp.tpg + pp.tsub;
  // the FFT tasks of a group are drawn `lag` slots after its FIR tasks: a little more than the tickets the
  // resident CTAs hold at any time, so the group is normally complete when its first FFT ticket is drawn
  int lag = h->pipe_lag > 0 ? h->pipe_lag : (int)((max_blocks * 5 / 4 + slot_len - 1) / slot_len) + 2;
  pp.lag = lag;
  pp.need_next = 0;
  for (int ph = 0; ph < prm.os; ph++) pp.need_next |= (int)(((prm.row_base + ph) / prm.os) & 1);
  pp.total = (long long)(pp.ngroups_fir + lag) * slot_len;
  const size_t ctrl_bytes = 16 + sizeof(int) * (size_t)(pp.ngroups_fir + 2);
  CHZ_CUDA(h->pipe_ctrl.reserve(ctrl_bytes));
  CHZ_CUDA(cudaMemsetAsync(h->pipe_ctrl.p, 0, ctrl_bytes, st));
  pp.ticket = (unsigned long long*)h->pipe_ctrl.p;
  pp.done = (int*)((char*)h->pipe_ctrl.p + 16);
  long long blocks = pp.total < max_blocks ? pp.total : max_blocks;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, 256, smem, st>>>(prm, pp);
  h->launches++;
  CHZ_CUDA(cudaGetLastError());
  return CHZ_OK;
}

template <bool IN16>
static int launch_pipe_dispatch(::chz* h, const ChanParams& prm, cudaStream_t st) {
#define CHZ_PIPE_P(MV)                                                 \
  switch (h->P) {                                                      \
    case 8: return launch_pipe<MV, 8, IN16>(h, prm, st);               \
    case 12: return launch_pipe<MV, 12, IN16>(h, prm, st);             \
    case 16: return launch_pipe<MV, 16, IN16>(h, prm, st);             \
    default: return 1;                                                 \
  }
  switch (h->M) {
    case 1024: CHZ_PIPE_P(1024)
    case 2048: CHZ_PIPE_P(2048)
    case 4096: CHZ_PIPE_P(4096)
    default: return 1;
  }
#undef CHZ_PIPE_P
}

bool pipe_available(const ::chz* h) {
  return (h->M == 1024 || h->M == 2048 || h->M == 4096) && (h->P == 8 || h->P == 12 || h->P == 16);
}


int launch_dit2_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st) {
  return in16 ? launch_dit2_dispatch<true>(h, prm, st) : launch_dit2_dispatch<false>(h, prm, st);
}
int launch_ws_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st) {
  return in16 ? launch_ws_dispatch<true>(h, prm, st) : launch_ws_dispatch<false>(h, prm, st);
}
int launch_cluster_any(::chz* h, const ChanParams& prm, bool in16, int path, cudaStream_t st) {
  switch (path) {
    case 7: return in16 ? launch_cluster_dispatch<true, 256, false>(h, prm, st) : launch_cluster_dispatch<false, 256, false>(h, prm, st);
    case 8: return in16 ? launch_cluster_dispatch<true, 256, true>(h, prm, st) : launch_cluster_dispatch<false, 256, true>(h, prm, st);
    case 9: return in16 ? launch_cluster_dispatch<true, 512, true>(h, prm, st) : launch_cluster_dispatch<false, 512, true>(h, prm, st);
    default: return in16 ? launch_cluster_dispatch<true, 512, false>(h, prm, st) : launch_cluster_dispatch<false, 512, false>(h, prm, st);
  }
}
int launch_dsm_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st) {
  return in16 ? launch_dsm_dispatch<true>(h, prm, st) : launch_dsm_dispatch<false>(h, prm, st);
}
int launch_pipe_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st) {
  return in16 ? launch_pipe_dispatch<true>(h, prm, st) : launch_pipe_dispatch<false>(h, prm, st);
}

}  // namespace chzi

#else   // default build: the experiment kernels are not compiled; forcing one of their paths is CHZ_EINVAL

namespace chzi {
struct ChanParams;
bool dit2_available(const ::chz*) { return false; }
bool ws_available(const ::chz*) { return false; }
bool cluster_available(const ::chz*, int) { return false; }
bool dsm_available(const ::chz*) { return false; }
bool pipe_available(const ::chz*) { return false; }
int launch_dit2_any(::chz*, const ChanParams&, bool, cudaStream_t) { return 1; }
int launch_ws_any(::chz*, const ChanParams&, bool, cudaStream_t) { return 1; }
int launch_cluster_any(::chz*, const ChanParams&, bool, int, cudaStream_t) { return 1; }
int launch_dsm_any(::chz*, const ChanParams&, bool, cudaStream_t) { return 1; }
int launch_pipe_any(::chz*, const ChanParams&, bool, cudaStream_t) { return 1; }
}  // namespace chzi

#endif
