/*
 * libchannelizer — C ABI of the B200-native polyphase channelizer + channelized PDW extractor.
 *
 * The reference (cwozny/sdr_channelizer) has no FFI for this path: its callers are MATLAB scripts.
 * Each entry point below states the reference lines it replaces (paths relative to the reference
 * root).  All functions return 0 (CHZ_OK) or a negative CHZ_E* code, in the status==0 /
 * bladerf_strerror() idiom of the reference's own C++ tools (cpp/blade_record_iq_12bit.cpp:54-59).
 * No exception crosses this boundary.  Handles are opaque and owned by the library; every buffer
 * passed in or out is caller-owned.  A handle is bound to the CUDA device that was current when it
 * was created and is NOT thread-safe; distinct handles may be used concurrently.
 *
 * There is no CPU fallback: every compute entry point fails with CHZ_ENODEVICE when no sm_100
 * device is usable.
 */
#ifndef CHANNELIZER_H
#define CHANNELIZER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHZ_ABI_VERSION 1

/* ---- error codes ---------------------------------------------------------------------------- */
#define CHZ_OK            0
#define CHZ_EINVAL       (-1)   /* bad argument (NULL, M not supported, ntaps % M != 0, ...)        */
#define CHZ_EIO          (-2)   /* open/read/mmap failure                                          */
#define CHZ_EFORMAT      (-3)   /* unknown magic ("Unsupported endianness", convert_my_iq_to_mat.m:55-56) */
#define CHZ_EBITWIDTH    (-4)   /* bitWidth outside (0,16]  ("Unsupported bit width", :96-97)       */
#define CHZ_ESIZE        (-5)   /* payload length != numSamples (the assert at :102)               */
#define CHZ_ENOMEM       (-6)   /* host or device allocation failed                                */
#define CHZ_ECUDA        (-7)   /* a CUDA call or kernel failed; see chz_last_cuda_error()         */
#define CHZ_ENODEVICE    (-8)   /* no usable sm_100 GPU (there is no CPU path)                     */
#define CHZ_ECAPACITY    (-9)   /* caller's output buffer too small; required size is returned     */
#define CHZ_ESTATE       (-10)  /* call not valid in the handle's current state                    */

const char* chz_strerror(int code);
/* Text of the last CUDA error seen by this thread's calls into the library ("" if none). */
const char* chz_last_cuda_error(void);
int chz_abi_version(void);

/* ---- R1: I/Q recording reader ---------------------------------------------------------------
 * Replaces matlab/convert_my_iq_to_mat.m:38-102 (the only parser in the reference) for the header
 * struct written by cpp/IqPacket.h:9-25 (format 3/2, 112 bytes) and by
 * matlab/generate_training_iq.m:107-125 (format 1, 104 bytes).
 */
typedef struct chz_iq chz_iq_t;

typedef struct chz_iq_info {
  uint32_t magic;              /* "endianness" word: 0x01010101 / 0x02020202 / 0x03030303 / 0      */
  uint32_t format;             /* 1, 2 or 3 (magic 0 -> 2, convert_my_iq_to_mat.m:43-45)           */
  uint32_t header_bytes;       /* 104 (format 1) or 112                                            */
  uint32_t link_speed;
  uint64_t fc_hz;              /* u32 on disk in format 1 (:63-65), u64 otherwise (:66-67)          */
  uint32_t bw_hz;
  uint32_t fs_sps;
  double   gain_db;            /* f32 on disk in format 3 (:73-74), u32 before (:75-77)             */
  uint32_t num_samples;        /* complex samples                                                   */
  uint32_t bit_width;          /* 1..16; <=8 -> int8 pairs, else int16 pairs (:92-98)               */
  uint32_t spare0;             /* absent (0) in format 1 (:82-84)                                   */
  uint32_t bytes_per_sample;   /* 2 or 4 (per complex sample)                                       */
  char board_name[17];         /* the four char[16] fields, NUL-stripped as strip(...,char(0))      */
  char serial_number[17];
  char fpga_version[17];
  char fw_version[17];
  double sample_start_time;    /* seconds since the Unix epoch (UTC)                                */
  uint64_t payload_offset;     /* == header_bytes                                                   */
  uint64_t payload_bytes;      /* num_samples * bytes_per_sample                                    */
} chz_iq_info_t;

/* Parse + validate + mmap.  *out may be NULL if only `info` is wanted (then nothing stays open). */
int chz_open_iq(const char* path, chz_iq_t** out, chz_iq_info_t* info);
/* Interleaved I,Q payload (int8 or int16, little endian), valid until chz_close_iq(). */
const void* chz_iq_payload(const chz_iq_t* f);
int chz_close_iq(chz_iq_t* f);
/* Writer used by the CLI/tests to make recordings the reference tools would have written
 * (cpp/blade_record_iq_12bit.cpp:320-323: header struct then raw payload).  format 1, 2 or 3. */
int chz_write_iq(const char* path, const chz_iq_info_t* info, const void* payload);

/* ---- R4: prototype filter -------------------------------------------------------------------
 * dsp.Channelizer(M) defaults (matlab/create_pdws_channelized.m:33): 12 taps per band, 80 dB.
 * Kaiser-windowed sinc, cutoff fs/(2M), unity DC gain.  `taps` receives M*taps_per_band floats. */
int chz_design_prototype(uint32_t M, uint32_t taps_per_band, double stopband_atten_db, float* taps);

/* ---- R2+R5: channelizer ---------------------------------------------------------------------
 * chz_create   <- channelizer = dsp.Channelizer(M)                 create_pdws_channelized.m:33
 * chz_process  <- iq = double(iq)/2^(bitWidth-1); iq = channelizer(iq)          :35-38, :57
 * chz_reset    <- reset(channelizer) / a fresh object per file                  :33 (inside loop)
 */
typedef struct chz chz_t;
typedef struct chz_cf32 { float re, im; } chz_cf32;

/* M: 1..4096.  M = 1 with the single tap 1.0 is the identity "channelizer" (unpack only), which
 * makes chz_pdws the wideband extractor of matlab/create_pdws.m.  Powers of two >= 8 and the reference's own
 * M = fs*1e-6 = 56 (create_pdws_channelized.m:31) and 560 run the tuned fused kernels; any other M with prime
 * factors <= 7 runs a split path with a run-time mixed-radix FFT; the rest a functional O(M^2) DFT path.
 * ntaps: positive multiple of M, ntaps/M <= 32; oversample 1 (D = M, critically sampled) or 2
 * (D = M/2, M even).  taps == NULL -> default prototype (12*M taps, 80 dB;
 * ntaps is then ignored).  Taps are copied. */
int chz_create(uint32_t M, const float* taps, uint32_t ntaps, uint32_t oversample, chz_t** out);
void chz_destroy(chz_t* h);
int chz_reset(chz_t* h);

/* Launch all work of this handle on the given cudaStream_t.  Until this is called the handle uses
 * its own non-blocking stream.  NULL means CUDA's default stream (as for any cudaStream_t);
 * (void*)-1 selects the handle's own stream again. */
int chz_set_stream(chz_t* h, void* cuda_stream);

/* Options (chz_set_option) */
#define CHZ_OPT_RETAIN        1  /* 1: chz_process also keeps its output rows on the device (a store that grows until chz_reset) so that
                                    chz_pdws can run over them; 0 (default): rows only go to `out` -- frame-by-frame streaming in the
                                    style of channelizer_example.m:50-56 then holds no memory beyond the FIR history */
#define CHZ_OPT_CHUNK_ROWS    2  /* host path: rows per pipelined H2D/compute/D2H chunk (0 = auto) */
#define CHZ_OPT_FORCE_PATH    3  /* kernel family: 0 auto (default), 1 fused kernel (M <= 560), 2 split FIR + row-FFT kernels,
                                    11 fused ring kernel (M = 1024: TMA-fed raw-sample ring, in-place FFT).  3-10 are the
                                    round-1 large-M / warp-specialisation experiments, present only in a `make EXPERIMENTS=1`
                                    build.  A path that is not built, or has no kernel for the handle's
                                    (M, taps), is refused with CHZ_EINVAL. */
#define CHZ_OPT_PDW_EVENT_PATH 4  /* 1: chz_pdws* always uses the edge-event path (events to the host, radix sort, pairing, second launch
                                    for the statistics) that the single-synchronisation extractor falls back to when a sample sits
                                    exactly on a representable threshold; 0 (default): automatic.  For A/B runs and tests. */
int chz_set_option(chz_t* h, int opt, int64_t value);

uint32_t chz_num_channels(const chz_t* h);
uint32_t chz_num_taps(const chz_t* h);
uint32_t chz_decimation(const chz_t* h);
int chz_get_taps(const chz_t* h, float* taps, uint32_t cap);
/* Rows the next chz_process call with `nsamp` new samples will produce (frame semantics: one row
 * per complete frame of D samples, i.e. the reference's trim, create_pdws_channelized.m:52-54). */
uint64_t chz_rows_for(const chz_t* h, uint64_t nsamp);

/* Streaming, stateful like the System object (matlab/channelizer_example.m:50-56): FIR history and
 * a partial frame persist across calls.  `iq` = nsamp interleaved I,Q pairs, int8 when
 * 0 < bit_width <= 8 else int16; normalised by 2^(bit_width-1) exactly as :35-37.
 * Output: row-major [row][channel], natural FFT order (channel 0 = DC), *nrows rows of M values.
 * out_cap_rows < rows needed -> CHZ_ECAPACITY with *nrows = rows needed (nothing consumed).
 * Host-pointer variant: pipelined H2D / kernels / D2H inside the call.  out may be NULL when
 * CHZ_OPT_RETAIN is on (rows are only kept on the device for chz_pdws). */
int chz_process(chz_t* h, const void* iq, uint64_t nsamp, uint32_t bit_width,
                chz_cf32* out, uint64_t out_cap_rows, uint64_t* nrows);
/* Device-pointer variant (benchmarks, pipelines that keep data resident): asynchronous on the
 * handle's stream; *nrows is written before return.  `out_dev` rows are NOT copied into the
 * retained store; use chz_pdws_dev on them. */
int chz_process_dev(chz_t* h, const void* iq_dev, uint64_t nsamp, uint32_t bit_width,
                    chz_cf32* out_dev, uint64_t out_cap_rows, uint64_t* nrows);
int chz_synchronize(chz_t* h);

/* R7: centre frequency offset (Hz) of natural-order channel k: k*fs/M for k < M/2, (k-M)*fs/M
 * otherwise, i.e. centerFrequencies(channelizer,fs) after the fftshift at :60 reads
 * (c - M/2)*fs/M for shifted column c = (k + M/2) mod M. */
double chz_channel_freq(const chz_t* h, uint32_t k, double fs);

/* K1 on its own (R2; create_pdws_channelized.m:35-38): int8/int16 pairs -> complex fp32. */
int chz_unpack_dev(const void* iq_dev, uint64_t nsamp, uint32_t bit_width, chz_cf32* out_dev,
                   void* cuda_stream);
/* K3 on its own: y[r][k] = sum_p u[r][p] e^{+j 2 pi k p / M} for `nrows` rows (the FFT stage of
 * the channelizer; validated against cuFFT in tests).  In-place allowed. */
int chz_fft_rows_dev(chz_t* h, const chz_cf32* u_dev, chz_cf32* y_dev, uint64_t nrows);

/* ---- R6,R8-R11: channelized PDW extraction --------------------------------------------------
 * Replaces matlab/create_pdws_channelized.m:60-136.
 */
typedef struct chz_pdw_params {
  double snr_threshold_db;     /* 15  (:74)  threshold = median * 10^(snr/10)  (as written, :75)    */
  double sat_level;            /* 0.9999 (:130)                                                     */
  double fc_hz;                /* recording centre frequency (fc, :80)                              */
  double fs_sps;               /* INPUT sample rate; the decimated rate fs/D is derived (:62)       */
  double t0;                   /* sampleStartTime (:98)                                             */
  uint32_t reproduce_phase_bug;/* 1: phase taken from shifted column 1 for every bin, as :114 does  */
  uint32_t use_trailing_threshold; /* 1: hysteresis, trailing edge at trailing_snr_threshold_db         */
  double trailing_snr_threshold_db;/* wideband script matlab/create_pdws.m:45-47: 18 dB up, 3 dB down   */
} chz_pdw_params_t;

typedef struct chz_pdw {
  double toa_s;                /* toa_row/fs_dec + t0, toa_row 1-based (:98)                        */
  double pw_s;                 /* (end_row - toa_row)/fs_dec (:110)                                 */
  double freq_hz;              /* fc + binFreq + fs_dec*median(wrapped phase diff)/360 (:114-122)   */
  double amp;                  /* median |y| over [toa_row, end_row] inclusive (:101)               */
  double snr_db;               /* 10*log10(amp/noise_floor) (:105)                                  */
  double noise_floor;          /* per-channel median |y| over the whole run (:73)                   */
  uint32_t channel;            /* shifted (centred) column index, 0-based: (k + M/2) mod M (:60)    */
  uint32_t channel_natural;    /* natural FFT-order channel k                                       */
  uint64_t toa_row;            /* 1-based row of the leading edge (:90)                             */
  uint64_t end_row;            /* 1-based row of the trailing edge (:94)                            */
  uint32_t saturated;          /* :130-132                                                          */
  uint32_t reserved;
} chz_pdw_t;

/* PDWs over every row retained since the last reset, ordered as the reference emits them:
 * shifted channel ascending, then time (:79,85).  If cap is too small returns CHZ_ECAPACITY and
 * *n = required count. */
int chz_pdws(chz_t* h, const chz_pdw_params_t* params, chz_pdw_t* out, uint64_t cap, uint64_t* n);
/* Same over a caller-owned device matrix y_dev[nrows][M] (natural channel order). */
int chz_pdws_dev(chz_t* h, const chz_pdw_params_t* params, const chz_cf32* y_dev, uint64_t nrows,
                 chz_pdw_t* out, uint64_t cap, uint64_t* n);
/* Records of the last chz_pdws* run (cached in the handle), without recomputing: use after a
 * CHZ_ECAPACITY answer.  *n = record count. */
int chz_pdws_fetch(const chz_t* h, chz_pdw_t* out, uint64_t cap, uint64_t* n);
/* Per-channel noise floor (natural order, M doubles) of the last chz_pdws* call. */
int chz_pdw_noise_floor(const chz_t* h, double* nf, uint32_t cap);

/* ---- time-sharded PDW extraction (one shard of the recording per GPU) --------------------------
 * create_pdws_channelized.m:73 takes the noise floor as the median over the WHOLE file, and a pulse may
 * straddle a shard boundary, so the extractor is exposed in stages; the host (one process per GPU,
 * sdr_channelizer_b200/sharding.py: create_pdws_sharded) runs them in lock-step and does the three small
 * exchanges in between (NCCL through torch.distributed):
 *   for pass = 0, 1, 2:  chz_pdw_shard_hist_dev -> all-reduce(sum) of the table -> chz_pdw_shard_select
 *   chz_pdw_shard_thresholds                       (identical noise floor / thresholds on every rank)
 *   chz_pdw_shard_exit_state_dev -> all-gather -> each rank folds the codes of the ranks before it
 *   chz_pdw_shard_detect_dev(entry state)          (events carry rows of the whole recording)
 *   all-gather events -> chz_pdw_pair_events       (same pulse list everywhere, reference order)
 *   chz_pdw_shard_records_dev on the pulses inside the shard; for a pulse that straddles shards the
 *   owner gathers the few column segments into a small matrix and calls it with ld = its width.
 * Results are identical to chz_pdws_dev over the stitched matrix (tests). */
typedef struct chz_pulse {
  uint64_t toa_row, end_row;   /* 1-based rows of the leading / trailing edge in the whole recording  */
  uint32_t channel_natural;    /* natural FFT-order channel of the pulse                              */
  uint32_t col, col_phase;     /* columns of the matrix handed to chz_pdw_shard_records_dev holding    */
                               /* the channel for |y| and for the phase (:114); = channel_natural when */
                               /* that matrix is the channel matrix itself                             */
  uint32_t reserved;
} chz_pulse_t;

/* Histogram of radix pass `pass` (0..2) over this shard's rows, accumulated into the handle's table
 * (uint32 [M][2][2048], device memory, returned in *hist_dev; *hist_words = its length).  The call
 * returns after the kernel has finished, so the table can be summed over ranks on any stream. */
int chz_pdw_shard_hist_dev(chz_t* h, const chz_cf32* y_dev, uint64_t nrows, int pass, uint32_t** hist_dev, uint64_t* hist_words);
/* Consume the (summed) table: fix the next bits of the two middle order statistics of `total_rows` values
 * per channel, then clear the table for the next pass. */
int chz_pdw_shard_select(chz_t* h, int pass, uint64_t total_rows);
/* After pass 2: noise floor (chz_pdw_noise_floor) and thresholds (:73-75). */
int chz_pdw_shard_thresholds(chz_t* h, const chz_pdw_params_t* params);
/* Instead of the three histogram/select passes and chz_pdw_shard_thresholds: take the per-channel noise floor
 * (natural order, M doubles) from the caller and derive the thresholds from it exactly as above.  For callers that
 * already know the floor (a calibration run, an earlier dwell) and for tests that feed the detector a double-
 * precision median, to tell threshold error from detection error. */
int chz_pdw_shard_set_noise_floor(chz_t* h, const chz_pdw_params_t* params, const double* noise_floor);
/* code[k] (host, M bytes, natural channels): state of the edge FSM after this shard's last row as a
 * function of the state it is entered with: 0 inactive, 1 active, 2 = entry state, 3 = entry state toggled. */
int chz_pdw_shard_exit_state_dev(chz_t* h, const chz_cf32* y_dev, uint64_t nrows, uint8_t* code);
/* Edge events of this shard.  entry: M bytes (host), FSM state per natural channel before the shard's
 * first row (NULL = inactive, :83).  event = (shifted channel << 40) | (1-based row of the recording << 1)
 * | (1 = trailing edge).  CHZ_ECAPACITY with *n = count when cap is too small. */
int chz_pdw_shard_detect_dev(chz_t* h, const chz_cf32* y_dev, uint64_t nrows, uint64_t row_offset, const uint8_t* entry,
                             uint64_t* events, uint64_t cap, uint64_t* n);
/* Host only: sort the events of ALL shards and pair them into pulses in the reference's output order
 * (shifted channel ascending, then time, :79,85); a pulse still open at the end is dropped (:135). */
int chz_pdw_pair_events(uint64_t* events, uint64_t n, uint32_t M, uint32_t reproduce_phase_bug, chz_pulse_t* pulses,
                        uint64_t cap, uint64_t* npulses);
/* Records (:97-128) of pulses whose rows all lie inside y_dev: a row-major matrix with leading dimension
 * ld whose first row is row row_offset + 1 of the recording. */
int chz_pdw_shard_records_dev(chz_t* h, const chz_pdw_params_t* params, const chz_cf32* y_dev, uint64_t ld,
                              uint64_t row_offset, const chz_pulse_t* pulses, uint64_t n, chz_pdw_t* out);

/* ---- event prediction from PDWs (host only) --------------------------------------------------------
 * The analysis the reference runs on the extractor's output (matlab/predict_event.m:125-138,
 * cpp/usrp_predict_event.cpp:28-52,348-373). */
/* Quadratic least-squares fit v ~ c0 + c1 t + c2 t^2 over n >= 3 points (Householder QR of [1 t t^2], i.e.
 * polyfit(pdw.toa, pdw.snr, 2), predict_event.m:125); *t_peak = -c1/(2 c2) (:128), *v_peak = the fit there
 * (:129); coef (optional) receives c0, c1, c2.  CHZ_EINVAL when the fit is rank deficient or has no peak. */
int chz_event_peak_time(const double* t, const double* v, uint64_t n, double* t_peak, double* v_peak, double* coef);
/* last event + median of the differences of successive event times (:133-135); one event: + fallback_interval
 * (:137).  upper_median != 0 takes element [size/2] of the sorted differences like the C++ tool
 * (usrp_predict_event.cpp:364-368) instead of MATLAB's median. */
int chz_next_event_time(const double* events, uint64_t n, double fallback_interval, int upper_median, double* next);

/* Device pointer and row count of the retained store (for callers that keep working on the GPU). */
int chz_retained(const chz_t* h, const chz_cf32** y_dev, uint64_t* nrows);
int chz_reserve_rows(chz_t* h, uint64_t nrows);

/* Number of kernels this handle has launched since creation (bench bookkeeping). */
uint64_t chz_kernel_launches(const chz_t* h);

/* Pinned host memory helpers for the host path. */
void* chz_alloc_host(uint64_t bytes);
void chz_free_host(void* p);

#ifdef __cplusplus
}
#endif
#endif /* CHANNELIZER_H */
