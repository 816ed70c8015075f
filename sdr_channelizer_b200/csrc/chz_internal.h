// Internal state of a channelizer handle (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "channelizer.h"
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost a few ns when no tool is attached

namespace chzi {

void set_cuda_error(cudaError_t e, const char* what, const char* file, int line);

#define CHZ_CUDA(call)                                                  \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) {                                           \
      ::chzi::set_cuda_error(e__, #call, __FILE__, __LINE__);            \
      return CHZ_ECUDA;                                                 \
    }                                                                   \
  } while (0)

}  // namespace chzi

namespace chzi {
// NVTX range per stage: "chz:channelize", "chz:pdw:median", "chz:pdw:detect", "chz:pdw:stats", ...
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Device scratch that lives with the handle and only ever grows: cudaMalloc/cudaFree on every call
// cost up to tens of milliseconds on a busy driver (measured), far more than the PDW kernels.
struct Scratch {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};
}  // namespace chzi

struct chz {
  int device = 0;
  int sm_count = 148;
  uint32_t M = 0, P = 0, L = 0, os = 1, D = 0;
  bool generic = false;                 // no compile-time radix plan for M (see mixed_np)
  int mixed_np = 0, mixed_r[12] = {0};  // generic M with prime factors <= 7: run-time mixed-radix plan; 0 = O(M^2) DFT
  int fir_bpb = 0;                      // register-window FIR block size for run-time M (0 = direct FIR kernel)
  std::vector<float> taps;              // prototype as given (unscaled), h[qM + p]
  float* d_taps[17] = {nullptr};        // per bit width: taps * 2^-(bw-1), uploaded on first use
  float2* d_tw = nullptr;               // inter-pass twiddles of the Stockham plan (per-pass layout, chz_create)
  float2* d_twn = nullptr;              // plain table e^{+j 2 pi i / M}, i < M (ring kernel)
  int ring_min_steps = 4, ring_unpack = 1, ring_variant = 0, ring_dbg = 0;   // ring kernel: 8-frame steps per CTA at least; tuning aids (CHZ_RING_* environment, read at chz_create)
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};

  // streaming state (stream indices count complex samples since reset)
  uint64_t consumed = 0, rows_done = 0;
  uint32_t bit_width = 0;               // of the current stream (0 = none yet)
  void* d_hist[2] = {nullptr, nullptr};
  int hist_cur = 0;
  uint64_t hist_base = 0, hist_len = 0, hist_cap = 0;   // samples

  // retained channel output (for chz_pdws)
  bool retain = false;                  // CHZ_OPT_RETAIN: off unless the caller wants chz_pdws over the processed rows
  bool poisoned = false;                // a chz_process call failed mid-pipeline: only chz_reset is valid
  float2* d_store = nullptr;
  uint64_t store_rows = 0, store_cap = 0;

  // cluster path: per-cluster L2-resident tile ring
  chzi::Scratch cluster_ring;

  // pipelined split path: ticket + per-group completion counters, geometry knobs
  chzi::Scratch pipe_ctrl;
  int pipe_span_rows = 64, pipe_lag = 0, pipe_blocks = 0;

  // split-path scratch (FIR output rows)
  float2* d_u = nullptr;
  uint64_t u_cap_rows = 0;

  // host-path staging
  void* d_in[2] = {nullptr, nullptr};
  uint64_t in_cap_bytes = 0;
  float2* d_out[2] = {nullptr, nullptr};
  uint64_t out_cap_rows = 0;
  int64_t chunk_rows = 0;

  int force_path = 0;
  bool split_overlap = false;                 // split path: run FFT(i) on a side stream under FIR(i+1)
  cudaEvent_t ev_fir[4] = {nullptr, nullptr, nullptr, nullptr}, ev_fft[4] = {nullptr, nullptr, nullptr, nullptr};
  uint64_t split_chunk_bytes = 1ull << 40;    // split path: FIR output bytes per launch pair (CHZ_SPLIT_CHUNK_MB)
  uint64_t launches = 0;

  // PDW scratch (histograms, select state, thresholds, edge events, pulse lists)
  chzi::Scratch pdw_hist, pdw_sel, pdw_thr, pdw_cnt, pdw_ev, pdw_pin, pdw_pout, pdw_code, pdw_nf, pdw_fast;
  uint64_t pdw_ev_cap = 1ull << 20;
  uint64_t pdw_pulse_cap = 1ull << 16;      // device-side pulse list of the one-GPU extractor (grows on demand)
  bool pdw_event_path = false;              // CHZ_OPT_PDW_EVENT_PATH
  void* pdw_stage_host = nullptr;           // pinned landing zone of its single device-to-host copy
  // CHZ_PDW_GRAPH=1: the extractor's eleven stream operations as one CUDA graph, re-used while the call's arguments
  // stay the same (a batch of equally sized files through one buffer).  Off by default: measured 0.234 ms per 100 ms
  // file against 0.231 ms launching them one by one -- the chain is bound by its kernels, not by its launches.
  cudaGraphExec_t pdw_graph = nullptr;
  unsigned char pdw_graph_key[160] = {0};
  int pdw_graph_kernels = 0;
  bool pdw_use_graph = false;
  size_t pdw_stage_bytes = 0;

  // PDW results of the last run
  std::vector<double> noise_floor;      // natural channel order
  std::vector<chz_pdw_t> pdws;
};

namespace chzi {
// pdw.cu
int pdw_extract(::chz* h, const chz_pdw_params_t* prm, const float2* y_dev, uint64_t nrows);
}  // namespace chzi
