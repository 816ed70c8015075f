"""ctypes binding of the CPU oracle (oracle/chz_oracle.cpp -> oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (sdr_channelizer_b200) never imports this module.
Channelizer parity is unpinned (dsp.Channelizer is closed source) — see the header of
chz_oracle.cpp; the header parser is pinned against the reference's own IqPacket struct.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "chz_oracle.cpp"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


class IqInfo(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("format", C.c_uint32), ("header_bytes", C.c_uint32),
                ("link_speed", C.c_uint32), ("fc_hz", C.c_uint64), ("bw_hz", C.c_uint32),
                ("fs_sps", C.c_uint32), ("gain_db", C.c_double), ("num_samples", C.c_uint32),
                ("bit_width", C.c_uint32), ("spare0", C.c_uint32), ("bytes_per_sample", C.c_uint32),
                ("board_name", C.c_char * 17), ("serial_number", C.c_char * 17),
                ("fpga_version", C.c_char * 17), ("fw_version", C.c_char * 17),
                ("sample_start_time", C.c_double), ("payload_offset", C.c_uint64),
                ("payload_bytes", C.c_uint64)]


class Pdw(C.Structure):
    _fields_ = [("toa_s", C.c_double), ("pw_s", C.c_double), ("freq_hz", C.c_double),
                ("amp", C.c_double), ("snr_db", C.c_double), ("noise_floor", C.c_double),
                ("channel", C.c_uint32), ("channel_natural", C.c_uint32),
                ("toa_row", C.c_uint64), ("end_row", C.c_uint64),
                ("saturated", C.c_uint32), ("reserved", C.c_uint32)]


class PdwParams(C.Structure):
    _fields_ = [("snr_threshold_db", C.c_double), ("sat_level", C.c_double), ("fc_hz", C.c_double),
                ("fs_sps", C.c_double), ("t0", C.c_double), ("reproduce_phase_bug", C.c_uint32),
                ("use_trailing_threshold", C.c_uint32), ("trailing_snr_threshold_db", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u64, u32, dp = C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_double)
        L.orc_parse_header.argtypes = [vp, u64, C.POINTER(IqInfo)]
        L.orc_parse_header.restype = C.c_int
        L.orc_unpack.argtypes = [vp, u64, u32, vp]
        L.orc_design_prototype.argtypes = [u32, u32, C.c_double, vp]
        L.orc_kaiser_beta.argtypes = [C.c_double]
        L.orc_kaiser_beta.restype = C.c_double
        L.orc_channelize_rows.argtypes = [vp, u64, u32, vp, u32, u32, u64, u64, vp]
        L.orc_channelize_rows.restype = u64
        L.orc_channelize_raw.argtypes = [vp, u64, u32, u32, vp, u32, u32, u64, u64, vp]
        L.orc_channelize_raw.restype = u64
        L.orc_channelize_direct.argtypes = [vp, u64, u32, vp, u32, u32, vp]
        L.orc_channelize_direct.restype = u64
        L.orc_center_frequencies.argtypes = [u32, C.c_double, vp]
        L.orc_pdws.argtypes = [vp, u64, u32, u32, C.POINTER(PdwParams), vp, u64, vp]
        L.orc_pdws.restype = u64
        L.orc_fsm_trace.argtypes = [vp, u64, C.c_double, vp, u64]
        L.orc_fsm_trace.restype = u64
        L.orc_median.argtypes = [vp, u64]
        L.orc_median.restype = C.c_double
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def parse_header(data: bytes):
    """-> (rc, IqInfo).  rc: 0 ok, -3 magic, -4 bit width, -5 size, -2 short."""
    info = IqInfo()
    buf = np.frombuffer(data, dtype=np.uint8)
    rc = lib().orc_parse_header(_ptr(buf), len(data), C.byref(info))
    return rc, info


def read_iq(path):
    """-> (IqInfo, payload ndarray [n,2] int8|int16).  Raises ValueError like the .m script errors."""
    with open(path, "rb") as f:
        data = f.read()
    rc, info = parse_header(data)
    if rc != 0:
        raise ValueError({-3: "Unsupported endianness", -4: "Unsupported bit width",
                          -5: "length(iq) != numSamples", -2: "short file"}[rc])
    dt = np.int8 if info.bit_width <= 8 else np.dtype("<i2")
    iq = np.frombuffer(data, dtype=dt, count=2 * info.num_samples, offset=info.payload_offset)
    return info, iq.reshape(-1, 2)


def unpack(iq, bit_width):
    """int8/int16 [n,2] -> complex128 [n]  (create_pdws_channelized.m:35-38)."""
    iq = np.ascontiguousarray(iq)
    n = iq.size // 2
    out = np.empty(n, dtype=np.complex128)
    lib().orc_unpack(_ptr(iq), n, bit_width, _ptr(out))
    return out


def design_prototype(M, taps_per_band=12, atten_db=80.0):
    h = np.empty(M * taps_per_band, dtype=np.float64)
    lib().orc_design_prototype(M, taps_per_band, atten_db, _ptr(h))
    return h


def channelize(x, M, taps, oversample=1, row0=0, nrows=0):
    """x complex128 [n] -> complex128 [rows, M], natural channel order."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    h = np.ascontiguousarray(taps, dtype=np.float64)
    D = M // oversample
    total = len(x) // D
    row0 = min(row0, total)
    nr = min(nrows, total - row0) if nrows else total - row0
    out = np.empty((nr, M), dtype=np.complex128)
    got = lib().orc_channelize_rows(_ptr(x), len(x), M, _ptr(h), len(h), oversample, row0, nr if nrows else 0, _ptr(out))
    assert got == nr
    return out


def channelize_raw(iq, bit_width, M, taps, oversample=1, row0=0, nrows=0):
    """Raw int8/int16 pairs -> complex128 [rows, M]; unpack fused (the timed CPU baseline)."""
    iq = np.ascontiguousarray(iq)
    n = iq.size // 2
    h = np.ascontiguousarray(taps, dtype=np.float64)
    D = M // oversample
    total = n // D
    row0 = min(row0, total)
    nr = min(nrows, total - row0) if nrows else total - row0
    out = np.empty((nr, M), dtype=np.complex128)
    got = lib().orc_channelize_raw(_ptr(iq), n, bit_width, M, _ptr(h), len(h), oversample, row0, nr if nrows else 0, _ptr(out))
    assert got == nr
    return out


def channelize_direct(x, M, taps, oversample=1):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    h = np.ascontiguousarray(taps, dtype=np.float64)
    rows = len(x) // (M // oversample)
    out = np.empty((rows, M), dtype=np.complex128)
    lib().orc_channelize_direct(_ptr(x), len(x), M, _ptr(h), len(h), oversample, _ptr(out))
    return out


def center_frequencies(M, fs):
    f = np.empty(M, dtype=np.float64)
    lib().orc_center_frequencies(M, fs, _ptr(f))
    return f


def pdws(y, D, snr_threshold_db=15.0, sat_level=0.9999, fc_hz=0.0, fs_sps=1.0, t0=0.0,
         reproduce_phase_bug=False, trailing_snr_threshold_db=None):
    """y complex128 [rows, M] natural order -> (list of Pdw, noise_floor[M] natural order)."""
    y = np.ascontiguousarray(y, dtype=np.complex128)
    rows, M = y.shape
    prm = PdwParams(snr_threshold_db, sat_level, fc_hz, fs_sps, t0, int(reproduce_phase_bug),
                    0 if trailing_snr_threshold_db is None else 1, float(trailing_snr_threshold_db or 0.0))
    nf = np.empty(M, dtype=np.float64)
    n = lib().orc_pdws(_ptr(y), rows, M, D, C.byref(prm), None, 0, _ptr(nf))
    arr = (Pdw * max(n, 1))()
    n2 = lib().orc_pdws(_ptr(y), rows, M, D, C.byref(prm), C.cast(arr, C.c_void_p), n, _ptr(nf))
    assert n2 == n
    return [arr[i] for i in range(n)], nf


def fsm_trace(mag, thr):
    mag = np.ascontiguousarray(mag, dtype=np.float64)
    out = np.zeros((len(mag), 2), dtype=np.uint64)
    n = lib().orc_fsm_trace(_ptr(mag), len(mag), float(thr), _ptr(out), len(mag))
    return [tuple(int(v) for v in out[i]) for i in range(n)]


def median(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_median(_ptr(v), len(v))


def num_threads():
    return lib().orc_num_threads()


# ---- event prediction (matlab/predict_event.m:125-138), numpy restatement ------------------------------
def event_peak_time(toa, snr):
    """p = polyfit(pdw.toa, pdw.snr, 2) (:125); t_max = -p(2)/(2*p(1)) (:128); y_max (:129)."""
    p = np.polyfit(np.asarray(toa, dtype=np.float64), np.asarray(snr, dtype=np.float64), 2)
    t_max = -p[1] / (2 * p[0])
    return t_max, p[0] * t_max ** 2 + p[1] * t_max + p[2]


def next_event_time(events, fallback_interval=4.61962892466417):
    """:133-138: median(diff(event)) + t_max, or t_max + 4.6196... for the first event."""
    e = np.asarray(events, dtype=np.float64)
    return e[-1] + (np.median(np.diff(e)) if len(e) > 1 else fallback_interval)
