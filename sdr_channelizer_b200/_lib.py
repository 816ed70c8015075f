"""ctypes binding of libchannelizer.so (include/channelizer.h).  No fallback: if the shared library is
missing or no B200 is usable the calls raise — nothing here computes on the CPU."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CHZ_LIB_PATH lets kernel experiments load an alternative build of the same ABI (never a different backend)
LIB_PATH = os.environ.get("CHZ_LIB_PATH") or os.path.join(_HERE, "libchannelizer.so")

CHZ_OK = 0
CHZ_EINVAL, CHZ_EIO, CHZ_EFORMAT, CHZ_EBITWIDTH, CHZ_ESIZE = -1, -2, -3, -4, -5
CHZ_ENOMEM, CHZ_ECUDA, CHZ_ENODEVICE, CHZ_ECAPACITY, CHZ_ESTATE = -6, -7, -8, -9, -10
CHZ_OPT_RETAIN, CHZ_OPT_CHUNK_ROWS, CHZ_OPT_FORCE_PATH, CHZ_OPT_PDW_EVENT_PATH = 1, 2, 3, 4

# every symbol include/channelizer.h declares (tests check the library exports them all)
EXPORTS = [
    "chz_strerror", "chz_last_cuda_error", "chz_abi_version",
    "chz_open_iq", "chz_iq_payload", "chz_close_iq", "chz_write_iq",
    "chz_design_prototype",
    "chz_create", "chz_destroy", "chz_reset", "chz_set_stream", "chz_set_option",
    "chz_num_channels", "chz_num_taps", "chz_decimation", "chz_get_taps", "chz_rows_for",
    "chz_process", "chz_process_dev", "chz_synchronize", "chz_channel_freq",
    "chz_unpack_dev", "chz_fft_rows_dev",
    "chz_pdws", "chz_pdws_dev", "chz_pdws_fetch", "chz_pdw_noise_floor", "chz_retained", "chz_reserve_rows",
    "chz_kernel_launches", "chz_alloc_host", "chz_free_host",
    "chz_pdw_shard_hist_dev", "chz_pdw_shard_select", "chz_pdw_shard_thresholds", "chz_pdw_shard_set_noise_floor", "chz_pdw_shard_exit_state_dev",
    "chz_pdw_shard_detect_dev", "chz_pdw_pair_events", "chz_pdw_shard_records_dev",
    "chz_event_peak_time", "chz_next_event_time",
]


class IqInfo(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("format", C.c_uint32), ("header_bytes", C.c_uint32),
                ("link_speed", C.c_uint32), ("fc_hz", C.c_uint64), ("bw_hz", C.c_uint32),
                ("fs_sps", C.c_uint32), ("gain_db", C.c_double), ("num_samples", C.c_uint32),
                ("bit_width", C.c_uint32), ("spare0", C.c_uint32), ("bytes_per_sample", C.c_uint32),
                ("board_name", C.c_char * 17), ("serial_number", C.c_char * 17),
                ("fpga_version", C.c_char * 17), ("fw_version", C.c_char * 17),
                ("sample_start_time", C.c_double), ("payload_offset", C.c_uint64),
                ("payload_bytes", C.c_uint64)]


class PdwParams(C.Structure):
    _fields_ = [("snr_threshold_db", C.c_double), ("sat_level", C.c_double), ("fc_hz", C.c_double),
                ("fs_sps", C.c_double), ("t0", C.c_double), ("reproduce_phase_bug", C.c_uint32),
                ("use_trailing_threshold", C.c_uint32), ("trailing_snr_threshold_db", C.c_double)]


class Pdw(C.Structure):
    _fields_ = [("toa_s", C.c_double), ("pw_s", C.c_double), ("freq_hz", C.c_double),
                ("amp", C.c_double), ("snr_db", C.c_double), ("noise_floor", C.c_double),
                ("channel", C.c_uint32), ("channel_natural", C.c_uint32),
                ("toa_row", C.c_uint64), ("end_row", C.c_uint64),
                ("saturated", C.c_uint32), ("reserved", C.c_uint32)]


class Pulse(C.Structure):
    _fields_ = [("toa_row", C.c_uint64), ("end_row", C.c_uint64), ("channel_natural", C.c_uint32),
                ("col", C.c_uint32), ("col_phase", C.c_uint32), ("reserved", C.c_uint32)]


class ChannelizerError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        msg = lib().chz_strerror(code).decode()
        if code == CHZ_ECUDA:
            msg += ": " + lib().chz_last_cuda_error().decode()
        super().__init__(f"{where}: {msg} ({code})" if where else f"{msg} ({code})")


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C sdr_channelizer_b200/csrc).  sdr_channelizer_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    pu64 = C.POINTER(C.c_uint64)
    L.chz_strerror.argtypes = [i32]; L.chz_strerror.restype = C.c_char_p
    L.chz_last_cuda_error.restype = C.c_char_p
    L.chz_abi_version.restype = i32
    L.chz_open_iq.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(IqInfo)]; L.chz_open_iq.restype = i32
    L.chz_iq_payload.argtypes = [vp]; L.chz_iq_payload.restype = vp
    L.chz_close_iq.argtypes = [vp]; L.chz_close_iq.restype = i32
    L.chz_write_iq.argtypes = [C.c_char_p, C.POINTER(IqInfo), vp]; L.chz_write_iq.restype = i32
    L.chz_design_prototype.argtypes = [u32, u32, C.c_double, vp]; L.chz_design_prototype.restype = i32
    L.chz_create.argtypes = [u32, vp, u32, u32, C.POINTER(vp)]; L.chz_create.restype = i32
    L.chz_destroy.argtypes = [vp]; L.chz_destroy.restype = None
    L.chz_reset.argtypes = [vp]; L.chz_reset.restype = i32
    L.chz_set_stream.argtypes = [vp, vp]; L.chz_set_stream.restype = i32
    L.chz_set_option.argtypes = [vp, i32, C.c_int64]; L.chz_set_option.restype = i32
    L.chz_num_channels.argtypes = [vp]; L.chz_num_channels.restype = u32
    L.chz_num_taps.argtypes = [vp]; L.chz_num_taps.restype = u32
    L.chz_decimation.argtypes = [vp]; L.chz_decimation.restype = u32
    L.chz_get_taps.argtypes = [vp, vp, u32]; L.chz_get_taps.restype = i32
    L.chz_rows_for.argtypes = [vp, u64]; L.chz_rows_for.restype = u64
    L.chz_process.argtypes = [vp, vp, u64, u32, vp, u64, pu64]; L.chz_process.restype = i32
    L.chz_process_dev.argtypes = [vp, vp, u64, u32, vp, u64, pu64]; L.chz_process_dev.restype = i32
    L.chz_synchronize.argtypes = [vp]; L.chz_synchronize.restype = i32
    L.chz_channel_freq.argtypes = [vp, u32, C.c_double]; L.chz_channel_freq.restype = C.c_double
    L.chz_unpack_dev.argtypes = [vp, u64, u32, vp, vp]; L.chz_unpack_dev.restype = i32
    L.chz_fft_rows_dev.argtypes = [vp, vp, vp, u64]; L.chz_fft_rows_dev.restype = i32
    L.chz_pdws.argtypes = [vp, C.POINTER(PdwParams), vp, u64, pu64]; L.chz_pdws.restype = i32
    L.chz_pdws_dev.argtypes = [vp, C.POINTER(PdwParams), vp, u64, vp, u64, pu64]; L.chz_pdws_dev.restype = i32
    L.chz_pdws_fetch.argtypes = [vp, vp, u64, pu64]; L.chz_pdws_fetch.restype = i32
    L.chz_pdw_noise_floor.argtypes = [vp, vp, u32]; L.chz_pdw_noise_floor.restype = i32
    L.chz_retained.argtypes = [vp, C.POINTER(vp), pu64]; L.chz_retained.restype = i32
    L.chz_reserve_rows.argtypes = [vp, u64]; L.chz_reserve_rows.restype = i32
    L.chz_kernel_launches.argtypes = [vp]; L.chz_kernel_launches.restype = u64
    L.chz_pdw_shard_hist_dev.argtypes = [vp, vp, u64, i32, C.POINTER(vp), pu64]; L.chz_pdw_shard_hist_dev.restype = i32
    L.chz_pdw_shard_select.argtypes = [vp, i32, u64]; L.chz_pdw_shard_select.restype = i32
    L.chz_pdw_shard_thresholds.argtypes = [vp, C.POINTER(PdwParams)]; L.chz_pdw_shard_thresholds.restype = i32
    L.chz_pdw_shard_set_noise_floor.argtypes = [vp, C.POINTER(PdwParams), vp]; L.chz_pdw_shard_set_noise_floor.restype = i32
    L.chz_pdw_shard_exit_state_dev.argtypes = [vp, vp, u64, vp]; L.chz_pdw_shard_exit_state_dev.restype = i32
    L.chz_pdw_shard_detect_dev.argtypes = [vp, vp, u64, u64, vp, vp, u64, pu64]; L.chz_pdw_shard_detect_dev.restype = i32
    L.chz_pdw_pair_events.argtypes = [vp, u64, u32, u32, vp, u64, pu64]; L.chz_pdw_pair_events.restype = i32
    L.chz_pdw_shard_records_dev.argtypes = [vp, C.POINTER(PdwParams), vp, u64, u64, vp, u64, vp]; L.chz_pdw_shard_records_dev.restype = i32
    L.chz_event_peak_time.argtypes = [vp, vp, u64, C.POINTER(C.c_double), C.POINTER(C.c_double), vp]; L.chz_event_peak_time.restype = i32
    L.chz_next_event_time.argtypes = [vp, u64, C.c_double, i32, C.POINTER(C.c_double)]; L.chz_next_event_time.restype = i32
    L.chz_alloc_host.argtypes = [u64]; L.chz_alloc_host.restype = vp
    L.chz_free_host.argtypes = [vp]; L.chz_free_host.restype = None
    _lib = L
    return L


def check(rc, where=""):
    if rc != CHZ_OK:
        raise ChannelizerError(rc, where)
