"""Host-side mirror of the reference's interface for the hot path, on top of the C ABI.

The reference's "API" for this path is a pair of MATLAB scripts plus the toolbox object they call:

    convert_my_iq_to_mat.m:38-118   .iq -> variables iq, fs, fc, dur, bw, gain, bitWidth, sampleStartTime, ...
    channelizer = dsp.Channelizer(M); y = channelizer(iq); centerFrequencies(channelizer, fs)
                                     (create_pdws_channelized.m:33,42,57; channelizer_example.m:31,56,60)
    create_pdws_channelized.m        pdw.toa / pdw.freq / pdw.pw / pdw.snr / pdw.sat

so this module offers read_iq(), Channelizer (same construct / call / centerFrequencies / reset
verbs, same property names) and create_pdws_channelized().  All arithmetic runs in the CUDA library;
numpy is only the container for host buffers.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import ChannelizerError, IqInfo, Pdw, PdwParams, check, lib


# ------------------------------------------------------------------------------------------------
# R1  .iq reader / writer
# ------------------------------------------------------------------------------------------------
class IqRecording:
    """Variables convert_my_iq_to_mat.m:118 saves, under the same names."""

    def __init__(self, info: IqInfo, iq: np.ndarray):
        self.iq = iq                                  # [N, 2] int8|int16 (MATLAB holds it as [2, N])
        self.fs = float(info.fs_sps)
        self.fc = float(info.fc_hz)
        self.bw = float(info.bw_hz)
        self.gain = float(info.gain_db)
        self.bitWidth = int(info.bit_width)
        self.sampleStartTime = float(info.sample_start_time)
        self.linkSpeed = int(info.link_speed)
        self.boardName = info.board_name.decode(errors="replace")
        self.serialNo = info.serial_number.decode(errors="replace")
        self.fpgaVersion = info.fpga_version.decode(errors="replace")
        self.fwVersion = info.fw_version.decode(errors="replace")
        self.fileFormat = int(info.format)
        self.numSamples = int(info.num_samples)
        self.dur = self.numSamples / self.fs if self.fs else 0.0   # convert_my_iq_to_mat.m:106
        self.info = info


def read_iq(path) -> IqRecording:
    """Parse a recording exactly as convert_my_iq_to_mat.m:38-102 does.  Errors mirror the script's:
    unknown magic -> 'Unsupported endianness', bad bitWidth -> 'Unsupported bit width', payload
    length != numSamples -> the assert at :102."""
    L = lib()
    handle = C.c_void_p()
    info = IqInfo()
    check(L.chz_open_iq(os.fsencode(path), C.byref(handle), C.byref(info)), f"chz_open_iq({path})")
    try:
        dt = np.int8 if info.bit_width <= 8 else np.dtype("<i2")
        n = int(info.num_samples)
        ptr = L.chz_iq_payload(handle)
        if n:
            buf = (C.c_char * (n * info.bytes_per_sample)).from_address(ptr)
            iq = np.frombuffer(buf, dtype=dt).reshape(n, 2).copy()
        else:
            iq = np.empty((0, 2), dtype=dt)
    finally:
        L.chz_close_iq(handle)
    return IqRecording(info, iq)


def write_iq(path, iq, *, fs, fc=0, bw=0, gain=0.0, bitWidth=16, sampleStartTime=0.0, fileFormat=3,
             linkSpeed=0, boardName="", serialNo="", fpgaVersion="", fwVersion=""):
    """Write a recording the way the recorders do (header struct, then raw interleaved payload;
    cpp/blade_record_iq_12bit.cpp:320-323).  iq: [N, 2] int8 (bitWidth <= 8) or int16."""
    dt = np.int8 if bitWidth <= 8 else np.dtype("<i2")
    iq = np.ascontiguousarray(iq, dtype=dt).reshape(-1, 2)
    info = IqInfo()
    info.format = fileFormat
    info.link_speed = linkSpeed
    info.fc_hz = int(fc)
    info.bw_hz = int(bw)
    info.fs_sps = int(fs)
    info.gain_db = float(gain)
    info.num_samples = iq.shape[0]
    info.bit_width = bitWidth
    info.board_name = boardName.encode()[:16]
    info.serial_number = serialNo.encode()[:16]
    info.fpga_version = fpgaVersion.encode()[:16]
    info.fw_version = fwVersion.encode()[:16]
    info.sample_start_time = float(sampleStartTime)
    check(lib().chz_write_iq(os.fsencode(path), C.byref(info), iq.ctypes.data_as(C.c_void_p)), "chz_write_iq")


def design_prototype(M, NumTapsPerBand=12, StopbandAttenuation=80.0) -> np.ndarray:
    taps = np.empty(M * NumTapsPerBand, dtype=np.float32)
    check(lib().chz_design_prototype(M, NumTapsPerBand, float(StopbandAttenuation), taps.ctypes.data_as(C.c_void_p)),
          "chz_design_prototype")
    return taps


class PdwTable:
    """Read-only sequence of Pdw records over the ctypes array the library filled: len(), indexing, iteration
    and comparison with a list, without building one Python object per record up front (at a few hundred
    records per 100 ms file that conversion cost more than the GPU extraction itself)."""

    def __init__(self, arr, n):
        self._arr, self._n = arr, int(n)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._arr[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._arr[i]

    def __iter__(self):
        return (self._arr[i] for i in range(self._n))

    def __eq__(self, other):
        return len(other) == self._n and all(a is b or bytes(a) == bytes(b) for a, b in zip(self, other))

    def to_numpy(self):
        """Structured numpy view of the records (field names of chz_pdw_t)."""
        dt = np.dtype([(name, ctype) for name, ctype in Pdw._fields_])
        return np.frombuffer(self._arr, dtype=dt, count=self._n) if self._n else np.empty(0, dtype=dt)


# ------------------------------------------------------------------------------------------------
# R4-R7  dsp.Channelizer mirror
# ------------------------------------------------------------------------------------------------
class Channelizer:
    """channelizer = dsp.Channelizer(M)   (create_pdws_channelized.m:33)

    NumFrequencyBands / NumTapsPerBand / StopbandAttenuation / OversamplingRatio follow the toolbox
    property names and defaults (12 taps per band, 80 dB, critically sampled).  `taps` overrides the
    designed prototype (length must be a multiple of M).  Stateful like the System object: FIR
    history carries across calls (channelizer_example.m:50-56) until reset().
    retain=True additionally keeps every output row on the GPU (CHZ_OPT_RETAIN) so that pdws() can run
    over everything processed since reset(); the default keeps nothing, as a streaming caller expects.
    """

    def __init__(self, NumFrequencyBands=8, NumTapsPerBand=12, StopbandAttenuation=80.0, OversamplingRatio=1,
                 taps=None, retain=False):
        self.NumFrequencyBands = int(NumFrequencyBands)
        self.OversamplingRatio = int(OversamplingRatio)
        if taps is None:
            taps = design_prototype(self.NumFrequencyBands, NumTapsPerBand, StopbandAttenuation)
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        self.NumTapsPerBand = len(taps) // self.NumFrequencyBands
        self.StopbandAttenuation = float(StopbandAttenuation)
        self._h = C.c_void_p()
        check(lib().chz_create(self.NumFrequencyBands, taps.ctypes.data_as(C.c_void_p), len(taps),
                               self.OversamplingRatio, C.byref(self._h)), "chz_create")
        self._taps = taps
        self._retain = False
        if retain:
            self.retain(True)

    def retain(self, on=True):
        """Keep (or stop keeping) the output rows of later calls on the GPU for pdws()."""
        self.set_option(_lib.CHZ_OPT_RETAIN, 1 if on else 0)
        self._retain = bool(on)

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().chz_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(lib().chz_reset(self._h), "chz_reset")

    def release(self):      # System-object verb
        self.reset()

    # -- properties --------------------------------------------------------------------------
    @property
    def handle(self):
        return self._h

    @property
    def DecimationFactor(self):
        return lib().chz_decimation(self._h)

    def coeffs(self):
        """Prototype low-pass coefficients (toolbox: coeffs(channelizer))."""
        return self._taps.copy()

    def centerFrequencies(self, fs):
        """centerFrequencies(channelizer, fs) as the reference uses it: against the fftshift-ed columns
        (create_pdws_channelized.m:42,60,80) -> ascending (-M/2 .. M/2-1) * fs / M."""
        M = self.NumFrequencyBands
        nat = np.array([lib().chz_channel_freq(self._h, k, float(fs)) for k in range(M)])
        return np.fft.fftshift(nat)

    def set_stream(self, cuda_stream_ptr):
        """cudaStream_t as an integer (e.g. torch.cuda.current_stream().cuda_stream; 0 = default stream);
        None selects the handle's own non-blocking stream."""
        ptr = C.c_void_p(-1) if cuda_stream_ptr is None else C.c_void_p(int(cuda_stream_ptr))
        check(lib().chz_set_stream(self._h, ptr), "chz_set_stream")

    def set_option(self, opt, value):
        check(lib().chz_set_option(self._h, opt, int(value)), "chz_set_option")

    @property
    def kernel_launches(self):
        return int(lib().chz_kernel_launches(self._h))

    # -- processing --------------------------------------------------------------------------
    def rows_for(self, nsamp):
        return int(lib().chz_rows_for(self._h, int(nsamp)))

    def __call__(self, iq, bitWidth, out=None):
        """y = channelizer(iq): iq is the RAW recording payload ([N, 2] int8/int16, as read_iq
        returns it); normalisation by 2^(bitWidth-1) (create_pdws_channelized.m:35-38) happens on the
        GPU.  Returns complex64 [rows, M], natural FFT channel order (apply np.fft.fftshift(y, axes=1)
        for the reference's :60)."""
        want = np.int8 if bitWidth <= 8 else np.dtype("<i2")
        iq = np.ascontiguousarray(iq)
        if iq.dtype != want:
            raise TypeError(f"bitWidth {bitWidth} needs {want} samples, got {iq.dtype}")
        nsamp = iq.size // 2
        M = self.NumFrequencyBands
        rows = self.rows_for(nsamp)
        if out is None:
            out = np.empty((rows, M), dtype=np.complex64)
        n = C.c_uint64(0)
        check(lib().chz_process(self._h, iq.ctypes.data_as(C.c_void_p), nsamp, bitWidth,
                                out.ctypes.data_as(C.c_void_p), out.shape[0], C.byref(n)), "chz_process")
        return out[: n.value]

    def process_ptr(self, iq_ptr, nsamp, bitWidth, out_ptr, out_cap_rows, device=True):
        """Raw-pointer entry (device pointers by default): returns rows produced.  Asynchronous on the
        handle's stream for device pointers."""
        n = C.c_uint64(0)
        fn = lib().chz_process_dev if device else lib().chz_process
        check(fn(self._h, C.c_void_p(iq_ptr), int(nsamp), int(bitWidth), C.c_void_p(out_ptr), int(out_cap_rows),
                 C.byref(n)), "chz_process_dev" if device else "chz_process")
        return int(n.value)

    def synchronize(self):
        check(lib().chz_synchronize(self._h), "chz_synchronize")

    def fft_rows_ptr(self, u_ptr, y_ptr, nrows):
        check(lib().chz_fft_rows_dev(self._h, C.c_void_p(u_ptr), C.c_void_p(y_ptr), int(nrows)), "chz_fft_rows_dev")

    # -- PDWs ----------------------------------------------------------------------------------
    def _pdw_call(self, fn, params, *lead):
        L = lib()
        n = C.c_uint64(0)
        # one call in the common case; a fresh array per call (the caller keeps the table), but neither a fresh array
        # TYPE per call (capacities are powers of two, ctypes caches the types) nor a zero-filled one (numpy memory)
        cap = max(256, 1 << (2 * getattr(self, "_last_pdw_count", 0)).bit_length())
        arr = (Pdw * cap).from_buffer(np.empty(cap * C.sizeof(Pdw), dtype=np.uint8))
        rc = fn(self._h, C.byref(params), *lead, C.cast(arr, C.c_void_p), cap, C.byref(n))
        cnt = int(n.value)
        if rc == _lib.CHZ_ECAPACITY:   # the run cached its records in the handle; copy them out without recomputing
            arr = (Pdw * cnt)()
            check(L.chz_pdws_fetch(self._h, C.cast(arr, C.c_void_p), cnt, C.byref(n)), "chz_pdws_fetch")
        elif rc != _lib.CHZ_OK:
            check(rc, "chz_pdws")
        self._last_pdw_count = cnt
        nf = np.empty(self.NumFrequencyBands, dtype=np.float64)
        check(L.chz_pdw_noise_floor(self._h, nf.ctypes.data_as(C.c_void_p), len(nf)), "chz_pdw_noise_floor")
        return PdwTable(arr, cnt), nf

    def pdws(self, fs, fc=0.0, sampleStartTime=0.0, SNR_THRESHOLD=15.0, sat_level=0.9999,
             reproduce_phase_bug=False, TRAILING_EDGE_THRESHOLD=None):
        """PDWs over everything processed since reset (create_pdws_channelized.m:60-136); the handle must
        have been retaining its rows (retain=True / retain()).
        -> (list of Pdw records in the reference's order, noise floor per natural channel)."""
        if not self._retain:
            raise ChannelizerError(_lib.CHZ_ESTATE, "pdws() needs Channelizer(..., retain=True)")
        prm = self._params(fs, fc, sampleStartTime, SNR_THRESHOLD, sat_level, reproduce_phase_bug, TRAILING_EDGE_THRESHOLD)
        return self._pdw_call(lib().chz_pdws, prm)

    @staticmethod
    def _params(fs, fc, t0, snr, sat, bug, trailing):
        return PdwParams(float(snr), float(sat), float(fc), float(fs), float(t0), int(bool(bug)),
                         0 if trailing is None else 1, float(trailing or 0.0))

    def pdws_ptr(self, y_ptr, nrows, fs, fc=0.0, sampleStartTime=0.0, SNR_THRESHOLD=15.0, sat_level=0.9999,
                 reproduce_phase_bug=False, TRAILING_EDGE_THRESHOLD=None):
        prm = self._params(fs, fc, sampleStartTime, SNR_THRESHOLD, sat_level, reproduce_phase_bug, TRAILING_EDGE_THRESHOLD)
        return self._pdw_call(lib().chz_pdws_dev, prm, C.c_void_p(y_ptr), int(nrows))


def unpack_ptr(iq_ptr, nsamp, bitWidth, out_ptr, stream_ptr=0):
    """K1 alone on device pointers (create_pdws_channelized.m:35-38)."""
    check(lib().chz_unpack_dev(C.c_void_p(iq_ptr), int(nsamp), int(bitWidth), C.c_void_p(out_ptr),
                               C.c_void_p(stream_ptr or 0)), "chz_unpack_dev")


# ------------------------------------------------------------------------------------------------
# create_pdws_channelized.m as a function
# ------------------------------------------------------------------------------------------------
def create_pdws_channelized(recordings, M=None, NumTapsPerBand=12, SNR_THRESHOLD=15.0,
                            reproduce_phase_bug=False, taps=None):
    """The reference script as a function: for each recording (path or IqRecording) build a fresh
    channelizer with M = fs*1e-6 bands unless given (:31-33), channelize, extract PDWs and
    concatenate (:16-20,124-128).  Any M in [1, 4096] (tuned kernels for powers of two, 56 and 560).
    reproduce_phase_bug=True gives the frequencies the script as written emits (its :114 reads column 1's phase
    for every bin); the default is the intended per-bin phase.
    Returns dict(toa, freq, pw, snr, sat, amp, channel) of numpy arrays."""
    out = {k: [] for k in ("toa", "freq", "pw", "snr", "sat", "amp", "channel")}
    for rec in recordings:
        if not isinstance(rec, IqRecording):
            rec = read_iq(rec)
        m = int(M) if M else int(round(rec.fs * 1e-6))            # :31
        ch = Channelizer(m, NumTapsPerBand=NumTapsPerBand, taps=taps, retain=True)   # :33
        try:
            n = C.c_uint64(0)
            iq = np.ascontiguousarray(rec.iq)
            check(lib().chz_process(ch.handle, iq.ctypes.data_as(C.c_void_p), iq.shape[0], rec.bitWidth,
                                    None, 0, C.byref(n)), "chz_process")                     # :35-57
            recs, _ = ch.pdws(rec.fs, rec.fc, rec.sampleStartTime, SNR_THRESHOLD,
                              reproduce_phase_bug=reproduce_phase_bug)                       # :60-136
        finally:
            ch.close()
        for r in recs:
            out["toa"].append(r.toa_s); out["freq"].append(r.freq_hz); out["pw"].append(r.pw_s)
            out["snr"].append(r.snr_db); out["sat"].append(bool(r.saturated)); out["amp"].append(r.amp)
            out["channel"].append(r.channel)
    return {k: np.asarray(v) for k, v in out.items()}


def channelizer_example(rec, duration=5e-3, step_frames=100, numBands=None, NumTapsPerBand=12, taps=None):
    """The math of matlab/channelizer_example.m:18-61 (the surf / VideoWriter part is the caller's) as a generator of
    frames (f_MHz, t_s, zeroCenterOut):

        iq = iq'                                   conjugate transpose (:23)
        channelizer = dsp.Channelizer(fs*1e-6)     1 MHz bins (:29-31)
        for ii = 1 : 100*numBands : length(iq)     windows of 5 ms, stepped by 100 frames (:33-34,50-55)
            out = abs(channelizer(iq(ii : ii+samples-1)))   ONE stateful object: the FIR history of the previous --
                                                            overlapping -- window carries into the next call (:56)
            zeroCenterOut = fftshift(out, 2)       (:58)
            f = (fc - centerFrequencies(channelizer, fs))*1e-6;  t = ii/fs + (0:rows-1)*numBands/fs   (:60-61)

    The conjugate is not applied to the raw integers (-32768 has no int16 negative): for real taps
    channelizer(conj(x))[:, k] = conj(channelizer(x)[:, (M-k) mod M]), so abs() of the conjugated input is abs() of the
    recording's own channels in mirrored order.  That is exact in real arithmetic; in fp32 the two differ by rounding
    only (tested against a double-precision evaluation of the conjugated samples)."""
    if not isinstance(rec, IqRecording):
        rec = read_iq(rec)
    M = int(numBands) if numBands else int(round(rec.fs * 1e-6))                     # :29
    samples = int(round(duration * rec.fs))                                            # :34
    ch = Channelizer(M, NumTapsPerBand=NumTapsPerBand, taps=taps)                      # :31
    try:
        f = (rec.fc - ch.centerFrequencies(rec.fs)) * 1e-6                             # :60
        mirror = (-np.arange(M)) % M
        iq = np.ascontiguousarray(rec.iq)
        for ii in range(1, iq.shape[0] + 1, step_frames * M):                          # :50 (ii is 1-based as in the script)
            start, stop = ii, ii + samples - 1                                         # :52-53
            if stop > iq.shape[0]:                                                     # :55
                continue
            out = np.abs(ch(iq[start - 1:stop], rec.bitWidth))[:, mirror]              # :56 (with :23 folded in)
            zeroCenterOut = np.fft.fftshift(out, axes=1)                               # :58
            t = ii / rec.fs + np.arange(out.shape[0]) * M / rec.fs                     # :61
            yield f, t, zeroCenterOut
    finally:
        ch.close()


def create_pdws(recordings, SNR_THRESHOLD=18.0, TRAILING_EDGE_THRESHOLD=3.0):
    """matlab/create_pdws.m as a function: the wideband (un-channelized) extractor.  The raw stream is
    normalised (:29-32) by a one-channel identity "channelizer" (K1 alone) and fed to the same PDW
    kernels with hysteresis: leading edge at 18 dB over the median magnitude, trailing edge at 3 dB
    (:44-47,58,63).  Returns dict(toa, freq, pw, mag, snr, sat) like the script's pdw struct (:86-91)."""
    out = {k: [] for k in ("toa", "freq", "pw", "mag", "snr", "sat")}
    for rec in recordings:
        if not isinstance(rec, IqRecording):
            rec = read_iq(rec)
        ch = Channelizer(1, taps=np.ones(1, dtype=np.float32), retain=True)
        try:
            n = C.c_uint64(0)
            iq = np.ascontiguousarray(rec.iq)
            check(lib().chz_process(ch.handle, iq.ctypes.data_as(C.c_void_p), iq.shape[0], rec.bitWidth,
                                    None, 0, C.byref(n)), "chz_process")
            recs, _ = ch.pdws(rec.fs, rec.fc, rec.sampleStartTime, SNR_THRESHOLD,
                              TRAILING_EDGE_THRESHOLD=TRAILING_EDGE_THRESHOLD)
        finally:
            ch.close()
        for r in recs:
            out["toa"].append(r.toa_s); out["freq"].append(r.freq_hz); out["pw"].append(r.pw_s)
            out["mag"].append(r.amp); out["snr"].append(r.snr_db); out["sat"].append(bool(r.saturated))
    return {k: np.asarray(v) for k, v in out.items()}


def event_peak_time(toa, snr):
    """p = polyfit(toa, snr, 2); t_max = -p(2)/(2 p(1)); y_max = polyval(p, t_max)
    (matlab/predict_event.m:125-129, cpp/usrp_predict_event.cpp:28-52).  -> (t_max, y_max, (c0, c1, c2))."""
    t = np.ascontiguousarray(toa, dtype=np.float64)
    v = np.ascontiguousarray(snr, dtype=np.float64)
    tp, vp = C.c_double(0), C.c_double(0)
    coef = np.zeros(3, dtype=np.float64)
    check(lib().chz_event_peak_time(t.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), len(t), C.byref(tp),
                                    C.byref(vp), coef.ctypes.data_as(C.c_void_p)), "chz_event_peak_time")
    return tp.value, vp.value, tuple(coef)


def next_event_time(events, fallback_interval=4.61962892466417, upper_median=False):
    """median(diff(event)) + t_max, or t_max + the script's fixed interval for the first event
    (predict_event.m:133-138); upper_median=True is the C++ tool's median (usrp_predict_event.cpp:364-368)."""
    e = np.ascontiguousarray(events, dtype=np.float64)
    nxt = C.c_double(0)
    check(lib().chz_next_event_time(e.ctypes.data_as(C.c_void_p), len(e), float(fallback_interval), int(bool(upper_median)),
                                    C.byref(nxt)), "chz_next_event_time")
    return nxt.value


def predict_event(recordings, SNR_THRESHOLD=20.0, min_peak=0.9):
    """matlab/predict_event.m as a function.  For every recording whose normalised peak magnitude exceeds 0.9
    (:52) extract wideband PDWs with one 20 dB threshold over the median magnitude (:63-65,76-83; TOAs relative
    to the first recording's sampleStartTime, :88), fit SNR against TOA with a parabola and take its vertex as
    the event time (:125-131), then predict the next event from the median spacing (:133-138).
    Returns dict(event, next_event, y_max, pdws_per_file)."""
    out = {"event": [], "next_event": [], "y_max": [], "pdws_per_file": []}
    first = None
    for rec in recordings:
        if not isinstance(rec, IqRecording):
            rec = read_iq(rec)
        if first is None:
            first = rec.sampleStartTime                                   # :46-48
        iq = np.ascontiguousarray(rec.iq)
        full = float(2 ** (rec.bitWidth - 1))
        peak2 = int(np.max(iq[:, 0].astype(np.int64) ** 2 + iq[:, 1].astype(np.int64) ** 2)) if len(iq) else 0
        if not peak2 > (min_peak * full) ** 2:                             # :52  max(abs(iq)) > 0.9
            out["pdws_per_file"].append(0)
            continue
        ch = Channelizer(1, taps=np.ones(1, dtype=np.float32), retain=True)
        try:
            n = C.c_uint64(0)
            check(lib().chz_process(ch.handle, iq.ctypes.data_as(C.c_void_p), iq.shape[0], rec.bitWidth,
                                    None, 0, C.byref(n)), "chz_process")
            recs, _ = ch.pdws(rec.fs, rec.fc, rec.sampleStartTime - first, SNR_THRESHOLD)   # :63-123
        finally:
            ch.close()
        out["pdws_per_file"].append(len(recs))
        if len(recs) < 3:
            continue
        t_max, y_max, _ = event_peak_time([r.toa_s for r in recs], [r.snr_db for r in recs])   # :125-129
        out["event"].append(t_max)                                                               # :131
        out["y_max"].append(y_max)
        out["next_event"].append(next_event_time(out["event"]))                                  # :133-138
    return {k: np.asarray(v) for k, v in out.items()}


def stft(iq, bitWidth, fs, Window=None):
    """[s, f, t] = stft(iq, fs, 'Window', hamming(768), 'OverlapLength', 0) of matlab/spectrogram_my_iq.m:114 on
    the GPU.  A short-time Fourier transform without overlap IS a critically sampled filterbank with one tap
    per band: M = numel(Window) channels, prototype = the (time-reversed) window.  Segment m of the signal is
    row m+1 of the channelizer fed with one leading zero sample, up to the fixed phase factor
    e^{-j 2 pi k (M-1)/M} (the filterbank counts taps back from a row's newest sample and uses the e^{+j}
    kernel) which is applied here.  FFTLength is taken equal to the window length (the script leaves it at
    the toolbox default; the toolbox source is not available: unpinned).  Two-sided, centred frequency axis.
    -> (s complex64 [M, segments], f [M] Hz, t [segments] s)."""
    w = np.hamming(768) if Window is None else np.asarray(Window, dtype=np.float64)     # hamming(768): symmetric
    M = int(w.size)
    iq = np.ascontiguousarray(iq).reshape(-1, 2)
    nseg = iq.shape[0] // M
    ch = Channelizer(M, taps=np.ascontiguousarray(w[::-1], dtype=np.float32))
    try:
        x = np.zeros((1 + nseg * M + (M - 1), 2), dtype=iq.dtype)                       # one zero in front, M-1 behind
        x[1:1 + nseg * M] = iq[:nseg * M]
        y = ch(x, bitWidth)[1:1 + nseg]
    finally:
        ch.close()
    k = np.arange(M)
    s_nat = y * np.exp(-2j * np.pi * k * (M - 1) / M).astype(np.complex64)
    s = np.fft.fftshift(s_nat, axes=1).T                                               # 'centered' (the default range)
    f = (np.arange(M) - M // 2) * (fs / M)
    t = (np.arange(nseg) * M + M / 2) / fs
    return s, f, t


def spectrogram_my_iq(rec):
    """The math of matlab/spectrogram_my_iq.m:104-115 for one recording (path or IqRecording): normalise, STFT
    with a 768-point Hamming window and no overlap, power abs(s).^2 over (f + fc, t).  Plotting is the caller's.
    -> dict(power [768, segments], f_hz (absolute), t_s)."""
    if not isinstance(rec, IqRecording):
        rec = read_iq(rec)
    s, f, t = stft(rec.iq, rec.bitWidth, rec.fs)
    return {"power": np.abs(s) ** 2, "f_hz": f + rec.fc, "t_s": t}


__all__ = ["PdwTable", "channelizer_example", "IqRecording", "read_iq", "write_iq", "design_prototype", "Channelizer", "unpack_ptr",
           "create_pdws_channelized", "create_pdws", "predict_event", "event_peak_time", "next_event_time", "stft", "spectrogram_my_iq", "ChannelizerError"]
