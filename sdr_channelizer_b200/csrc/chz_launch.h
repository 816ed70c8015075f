// Launch plumbing shared by the translation units of libchannelizer (split so that make -j compiles the
// kernel families in parallel): span planning and the per-family launch entry points.
#pragma once
#include <cuda_runtime.h>

#include "chz_internal.h"

namespace chzi {
struct ChanParams;

struct LaunchPlan { dim3 grid; int span_rows; long long spans_per_phase; };
constexpr int kMaxDev = 64;   // function attributes (dynamic smem size) and occupancy are per device

LaunchPlan plan_spans(const ::chz* h, long long nrows, int P, int groups_per_block, int blocks_per_sm, int max_blocks_override = 0);

// every entry returns CHZ_OK, a CHZ_E* code, or 1 when there is no instantiation for the handle's (M, P)
int launch_fused_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st);      // chz_launch_fused.cu
int launch_dit2_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st);       // chz_launch_exp.cu ...
int launch_ws_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st);
int launch_cluster_any(::chz* h, const ChanParams& prm, bool in16, int path, cudaStream_t st);   // path 3, 7, 8 or 9
int launch_dsm_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st);
int launch_pipe_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st);
int launch_ring_any(::chz* h, const ChanParams& prm, bool in16, cudaStream_t st);      // chz_launch_ring.cu
bool ring_available(const ::chz* h);
bool dit2_available(const ::chz* h);
bool ws_available(const ::chz* h);
bool cluster_available(const ::chz* h, int tpc);
bool dsm_available(const ::chz* h);
bool pipe_available(const ::chz* h);
}  // namespace chzi
