"""PDW extraction (K4) through the C ABI against the oracle's transcription of
matlab/create_pdws_channelized.m:60-136.  north_star: record counts bit-exact, TOA/PW within +-1 row."""
import numpy as np
import pytest

import sdr_channelizer_b200 as pkg
from tests import synth
from tests.test_oracle import _pulse_matrix

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _pdws_on_matrix(y, fs, **kw):
    """Run K4 on a hand-built channel matrix (complex64 on the device)."""
    torch = _torch()
    M = y.shape[1]
    d = torch.from_numpy(np.ascontiguousarray(y.astype(np.complex64))).cuda()
    ch = pkg.Channelizer(M, NumTapsPerBand=8) if M > 1 else pkg.Channelizer(1, taps=np.ones(1, np.float32))
    ch.set_stream(torch.cuda.current_stream().cuda_stream)
    recs, nf = ch.pdws_ptr(d.data_ptr(), y.shape[0], fs, **kw)
    ch.close()
    return recs, nf


def _compare(recs, orecs, fs_dec, tol_rows=1):
    assert len(recs) == len(orecs), (len(recs), len(orecs))
    for a, b in zip(recs, orecs):
        assert (a.channel, a.channel_natural) == (b.channel, b.channel_natural)
        assert abs(int(a.toa_row) - int(b.toa_row)) <= tol_rows
        assert abs(int(a.end_row - a.toa_row) - int(b.end_row - b.toa_row)) <= tol_rows
        assert abs(a.toa_s - b.toa_s) <= tol_rows / fs_dec + 1e-12 and abs(a.pw_s - b.pw_s) <= tol_rows / fs_dec + 1e-12
        assert a.saturated == b.saturated
        assert abs(a.amp - b.amp) <= 1e-5 * b.amp and abs(a.snr_db - b.snr_db) <= 1e-3
        assert abs(a.freq_hz - b.freq_hz) <= 1e-4 * fs_dec, (a.freq_hz, b.freq_hz)


def test_known_answer_matrix(orc):
    M, fs, fc, t0 = 8, 8e6, 1e9, 1000.0
    y, k1, k2 = _pulse_matrix(M)
    y = y.astype(np.complex64)               # the oracle sees exactly the fp32 values the GPU gets
    recs, nf = _pdws_on_matrix(y, fs, fc=fc, sampleStartTime=t0)
    orecs, onf = orc.pdws(y.astype(np.complex128), M, fc_hz=fc, fs_sps=fs, t0=t0)
    assert np.allclose(nf, onf, rtol=2e-7)
    _compare(recs, orecs, fs / M, tol_rows=0)
    assert [(r.toa_row, r.end_row) for r in recs] == [(201, 232), (101, 152)]
    assert [r.saturated for r in recs] == [1, 0]


def test_fsm_equality_toggle_and_open_pulse(orc):
    """SNR_THRESHOLD = 0 dB and a median of 2^-6 make the threshold exactly representable, so samples
    EQUAL to it occur: >= opens and <= closes on the same value (:88,:94), i.e. equality toggles the
    state sample by sample.  Long runs of equal samples cross the detector's 64-row chunks."""
    M, rows = 8, 301
    y = np.full((rows, M), 2.0 ** -6, dtype=np.complex64)
    y[100:111, 3] = 0                       # below: forces inactive
    y[150:156, 3] = 1.5 * 2.0 ** -6         # above: forces active
    y[200:203, 5] = 3.0 * 2.0 ** -6
    y[rows - 1, 6] = 1.0                    # pulse still open at the end of the file: dropped (:135)
    recs, nf = _pdws_on_matrix(y, 8e6, SNR_THRESHOLD=0.0)
    orecs, onf = orc.pdws(y.astype(np.complex128), M, snr_threshold_db=0.0, fs_sps=8e6)
    assert np.all(nf == 2.0 ** -6) and np.all(onf == 2.0 ** -6)
    assert len(orecs) > 4 * (rows // 2 - 10)
    _compare(recs, orecs, 1e6, tol_rows=0)


def test_saturation_not_checked_on_edges(orc):
    M = 8
    y, k1, k2 = _pulse_matrix(M)
    y[210, k2] = 0.5
    y[200, k2] = 1.0
    y = y.astype(np.complex64)
    recs, _ = _pdws_on_matrix(y, 8e6)
    orecs, _ = orc.pdws(y.astype(np.complex128), M, fs_sps=8e6)
    _compare(recs, orecs, 1e6, tol_rows=0)
    assert [r.saturated for r in recs] == [0, 0]


def test_phase_bug_flag(orc):
    M = 8
    y, k1, k2 = _pulse_matrix(M)
    y[:, M // 2] = 0.01 * np.exp(1j * 0.5 * np.arange(y.shape[0]))
    y = y.astype(np.complex64)
    for flag in (False, True):
        recs, _ = _pdws_on_matrix(y, 8e6, reproduce_phase_bug=flag)
        orecs, _ = orc.pdws(y.astype(np.complex128), M, fs_sps=8e6, reproduce_phase_bug=flag)
        _compare(recs, orecs, 1e6, tol_rows=0)


@pytest.mark.parametrize("M,seed,hyst", [(256, 100, False), (64, 102, False), (8, 104, False), (1, 105, True)])
def test_single_sync_extractor_equals_event_path(M, seed, hyst):
    """The one-GPU extractor emits finished pulses on the device (k_detect<true> -> k_pulse_stats over the device-side
    list -> one copy, one synchronisation).  Its records must be byte-identical to those of the edge-event path
    (events to the host, radix sort, pairing, second launch) it replaced and still falls back to."""
    torch = _torch()
    if M == 1:      # the wideband recipe of test_wideband_create_pdws_script: pulses thousands of rows long
        fs = 10e6
        x, _ = synth.pulsed_complex(400_000, fs, seed=77, sigma=0.004, amp=0.6)
        iq = np.stack([np.clip(np.rint(x.real * 32768), -32768, 32767), np.clip(np.rint(x.imag * 32768), -32768, 32767)],
                      axis=1).astype(np.int16)
        bw = 16
    else:
        iq, bw, fs = synth.pulsed_int16(M * 9000, M=M, seed=seed)
    taps = pkg.design_prototype(M, 16) if M > 1 else np.ones(1, np.float32)
    kw = dict(SNR_THRESHOLD=18.0, TRAILING_EDGE_THRESHOLD=3.0) if hyst else {}
    out = []
    for event_path in (0, 1):
        ch = pkg.Channelizer(M, taps=taps, retain=True)
        ch.set_option(pkg.CHZ_OPT_PDW_EVENT_PATH, event_path)
        ch(iq, bw)
        recs, nf = ch.pdws(fs, 2.4e9, 17.0, **kw)
        out.append((b"".join(bytes(r) for r in recs), nf.copy(), len(recs)))
        ch.close()
    assert out[0][2] == out[1][2] and out[0][2] >= 3
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])


def test_cuda_graph_extractor_equals_direct_launches(monkeypatch):
    """CHZ_PDW_GRAPH=1 replays the extractor's stream operations as one CUDA graph while the call's arguments stay
    the same: the first call captures, the second replays, a call with other parameters captures again.  Records and
    noise floor must be byte-identical to the directly launched chain every time."""
    _torch()
    M = 64
    iq, bw, fs = synth.pulsed_int16(M * 9000, M=M, seed=102)
    taps = pkg.design_prototype(M, 16)
    out = {}
    for graph in ("0", "1"):
        monkeypatch.setenv("CHZ_PDW_GRAPH", graph)          # read when the handle is created
        ch = pkg.Channelizer(M, taps=taps, retain=True)
        ch(iq, bw)
        got = []
        for snr in (15.0, 15.0, 12.0, 15.0):
            recs, nf = ch.pdws(fs, 2.4e9, 17.0, SNR_THRESHOLD=snr)
            got.append((b"".join(bytes(r) for r in recs), nf.copy(), len(recs)))
        out[graph] = got
        ch.close()
    assert out["0"][0][2] >= 3 and out["0"][2][2] >= out["0"][0][2]
    for a, b in zip(out["0"], out["1"]):
        assert a[2] == b[2] and a[0] == b[0] and np.array_equal(a[1], b[1])
    assert out["1"][0][0] == out["1"][1][0] == out["1"][3][0]


@pytest.mark.parametrize("M,rows", [(56, 4001), (3, 9000), (2, 70001), (20, 1234), (16, 300), (4, 257), (48, 66000)])
def test_noise_floor_is_the_exact_median_for_any_channel_count(M, rows):
    """The histogram kernel gives a block 16 channels (lanes past the last channel idle; for M < 16 dividing 16 a
    16-lane row is 16/M matrix rows) and packs its counters in 16 bits, with the host keeping a block below 65 536
    rows.  The per-channel noise floor must be the median of the fp32 magnitudes for every such shape, with odd and
    even row counts; channels get different scales so that their selects differ."""
    rng = np.random.default_rng(M * 1000 + rows)
    scale = (1.0 + np.arange(M)) / M
    y = ((rng.standard_normal((rows, M)) + 1j * rng.standard_normal((rows, M))) * scale * 0.01).astype(np.complex64)
    recs, nf = _pdws_on_matrix(y, 1e6 * M)
    x, yy = y.real.astype(np.float64), y.imag.astype(np.float64)
    mag = np.sqrt((x * x + np.float32(1) * (yy * yy)).astype(np.float32)).astype(np.float32)   # |y| in fp32
    assert recs == []
    assert np.allclose(nf, np.median(mag.astype(np.float64), axis=0), rtol=3e-7, atol=0)


def test_no_pulses_and_empty():
    rng = np.random.default_rng(0)
    y = (rng.standard_normal((1000, 16)) + 1j * rng.standard_normal((1000, 16))).astype(np.complex64)
    recs, nf = _pdws_on_matrix(y, 16e6)
    assert recs == [] and np.allclose(nf, np.median(np.abs(y), axis=0), rtol=1e-6)


@pytest.mark.parametrize("M,P,seed", [(256, 16, 100), (256, 16, 101), (64, 12, 102), (1024, 16, 103), (8, 8, 104)])
def test_end_to_end_pulsed_recording(orc, M, P, seed, tmp_path):
    """configs[4]: generate_channelized_training_iq-style file -> channelizer -> PDWs, via the file
    reader and create_pdws_channelized(), against the oracle run on the same bytes."""
    _torch()
    n = M * 9000
    iq, bw, fs = synth.pulsed_int16(n, M=M, seed=seed)
    path = str(tmp_path / "pulsed.iq")
    pkg.write_iq(path, iq, fs=fs, fc=2.4e9, bitWidth=bw, sampleStartTime=1.7e9, fileFormat=1 if M <= 64 else 3)
    taps = pkg.design_prototype(M, P)
    rec = pkg.read_iq(path)
    ch = pkg.Channelizer(M, taps=taps, retain=True)
    y = ch(rec.iq, rec.bitWidth)
    recs, nf = ch.pdws(rec.fs, rec.fc, rec.sampleStartTime)
    ch.close()
    oy = orc.channelize_raw(rec.iq, rec.bitWidth, M, taps.astype(np.float64))
    assert synth.rel_rms(y, oy) <= 1e-5
    orecs, onf = orc.pdws(oy, M, fc_hz=rec.fc, fs_sps=rec.fs, t0=rec.sampleStartTime)
    assert len(orecs) >= 3, "generator should produce several pulses"
    assert np.allclose(nf, onf, rtol=1e-4)
    _compare(recs, orecs, rec.fs / M, tol_rows=1)
    # the script-level entry point gives the same table
    pdw = pkg.create_pdws_channelized([path], M=M, taps=taps)
    assert len(pdw["toa"]) == len(orecs)
    assert np.allclose(pdw["toa"], [r.toa_s for r in orecs], atol=1.0 / (rec.fs / M) + 1e-9)
    assert np.array_equal(pdw["sat"], [bool(r.saturated) for r in orecs])


def test_reference_script_defaults_m56(orc, tmp_path):
    """create_pdws_channelized.m as written: M = fs*1e-6 = 56 bands at 56 MS/s, dsp.Channelizer(M) with the
    toolbox defaults (12 taps per band, 80 dB), threshold 15 dB -- through the .iq reader and the script-level
    entry point, against the oracle on the same bytes."""
    _torch()
    M, fs = 56, 56e6
    n = M * 9000
    iq, bw, _ = synth.pulsed_int16(n, M=M, seed=321, fs=fs)
    path = str(tmp_path / "b200mini.iq")
    pkg.write_iq(path, iq, fs=fs, fc=5.8e9, bitWidth=bw, sampleStartTime=1.7e9, fileFormat=3, boardName="b200mini")
    pdw = pkg.create_pdws_channelized([path])                  # M and prototype from the file, like the script
    rec = pkg.read_iq(path)
    taps = orc.design_prototype(M, 12, 80.0)
    oy = orc.channelize_raw(rec.iq, rec.bitWidth, M, taps)
    orecs, _ = orc.pdws(oy, M, fc_hz=rec.fc, fs_sps=rec.fs, t0=rec.sampleStartTime)
    assert len(orecs) >= 3 and len(pdw["toa"]) == len(orecs)
    fs_dec = fs / M
    assert np.allclose(pdw["toa"], [r.toa_s for r in orecs], atol=1.0 / fs_dec + 1e-9)
    assert np.allclose(pdw["pw"], [r.pw_s for r in orecs], atol=1.0 / fs_dec + 1e-12)
    assert np.array_equal(pdw["channel"], [r.channel for r in orecs])
    assert np.allclose(pdw["freq"], [r.freq_hz for r in orecs], atol=1e-4 * fs_dec)


def test_wideband_create_pdws_script(orc, tmp_path):
    """matlab/create_pdws.m end to end: raw stream -> normalise (K1 via the one-channel identity
    channelizer) -> hysteresis detector (18 dB / 3 dB) -> per-pulse medians, against the oracle."""
    _torch()
    fs, n = 10e6, 400_000
    x, meta = synth.pulsed_complex(n, fs, seed=77, sigma=0.004, amp=0.6)
    iq = np.stack([np.clip(np.rint(x.real * 32768), -32768, 32767), np.clip(np.rint(x.imag * 32768), -32768, 32767)],
                  axis=1).astype(np.int16)
    path = str(tmp_path / "wide.iq")
    pkg.write_iq(path, iq, fs=fs, fc=915e6, bitWidth=16, sampleStartTime=100.0)
    pdw = pkg.create_pdws([path])
    y = orc.unpack(iq, 16).reshape(-1, 1)
    orecs, onf = orc.pdws(y, 1, snr_threshold_db=18.0, fc_hz=915e6, fs_sps=fs, t0=100.0, trailing_snr_threshold_db=3.0)
    assert len(orecs) >= 3 and len(pdw["toa"]) == len(orecs)
    assert np.allclose(pdw["toa"], [r.toa_s for r in orecs], atol=1.0 / fs + 1e-9)
    assert np.allclose(pdw["pw"], [r.pw_s for r in orecs], atol=1.0 / fs + 1e-12)
    assert np.allclose(pdw["mag"], [r.amp for r in orecs], rtol=1e-5)
    assert np.allclose(pdw["freq"], [r.freq_hz for r in orecs], atol=1e-4 * fs)
    assert np.array_equal(pdw["sat"], [bool(r.saturated) for r in orecs])
    # hysteresis known answer on the device matrix path
    t = np.full(2000, 0.01 + 0j, dtype=np.complex64)
    t[500:600] = 0.9; t[600:650] = 0.05; t[650] = 0.0101; t[900:905] = 0.8
    recs, _ = _pdws_on_matrix(t.reshape(-1, 1), 1e6, SNR_THRESHOLD=18.0, TRAILING_EDGE_THRESHOLD=3.0)
    assert [(r.toa_row, r.end_row) for r in recs] == [(501, 651), (901, 906)]


# ---- time-sharded extraction: distributed median + boundary stitching ---------------------
def _sharded_on_one_gpu(y, bounds, fs, **kw):
    """create_pdws_sharded with one thread per shard (each with its own handle and stream) on one GPU;
    the exchanges go through ThreadComm instead of NCCL, everything else is the multi-GPU code path."""
    import threading
    torch = _torch()
    from sdr_channelizer_b200.sharding import PdwShard, ThreadComm, create_pdws_sharded
    M = y.shape[1]
    d = torch.from_numpy(np.ascontiguousarray(y.astype(np.complex64))).cuda()
    torch.cuda.synchronize()
    world = len(bounds) - 1
    comms = ThreadComm.make(world)
    results, errors = [None] * world, []

    def run(r):
        try:
            ch = pkg.Channelizer(M, NumTapsPerBand=8) if M > 1 else pkg.Channelizer(1, taps=np.ones(1, np.float32))
            a, b = bounds[r], bounds[r + 1]
            shard = PdwShard(ch, d.data_ptr() + a * M * 8, b - a, a, y.shape[0], fs, **kw)
            results[r] = create_pdws_sharded(shard, comms[r])
            ch.close()
        except Exception as e:   # surface the failure instead of dead-locking the other threads' barriers
            errors.append(e)
            comms[r]._sh["barrier"].abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    return results


def _same_records(a, b):
    assert len(a) == len(b)
    for x, z in zip(a, b):
        assert bytes(x) == bytes(z), ((x.toa_row, x.end_row, x.channel, x.amp, x.freq_hz), (z.toa_row, z.end_row, z.channel, z.amp, z.freq_hz))


@pytest.mark.parametrize("bounds", [[0, 200, 400, 600], [0, 199, 600], [0, 1, 2, 600], [0, 64, 128, 192, 256, 320, 384, 600],
                                    [0, 0, 300, 300, 600, 600]])      # ranks that hold no rows at all
def test_sharded_pdws_identical_to_one_gpu(bounds):
    """Every boundary case (tests/test_sharding.py::_pdw_matrix): records and noise floor of the sharded
    extractor are byte-identical to chz_pdws_dev on the whole matrix, on every rank."""
    from tests.test_sharding import _pdw_matrix
    y = _pdw_matrix()
    kw = dict(fc=1e9, sampleStartTime=3.0)
    whole, nf = _pdws_on_matrix(y, 8e6, **kw)
    assert len(whole) == 7
    for recs, nfs in _sharded_on_one_gpu(y, bounds, 8e6, **kw):
        _same_records(recs, whole)
        assert np.array_equal(nfs, nf)


def test_sharded_pdws_equality_toggle_across_boundaries():
    """Threshold exactly representable (0 dB over a median of 2^-6): samples equal to it toggle the FSM
    (:88,:94), so a shard's exit state is 'entry toggled n times' and must be folded across shards."""
    M, rows = 8, 301
    y = np.full((rows, M), 2.0 ** -6, dtype=np.complex64)
    y[10:40, 1] = 0.5; y[100:230, 2] = 0.25; y[::2, 3] = 0.001; y[5, 4] = 0.7
    whole, nf = _pdws_on_matrix(y, 8e6, SNR_THRESHOLD=0.0)
    assert len(whole) > 100
    for bounds in ([0, 100, 200, 301], [0, 151, 301], [0, 7, 301]):
        for recs, nfs in _sharded_on_one_gpu(y, bounds, 8e6, SNR_THRESHOLD=0.0):
            _same_records(recs, whole)


def test_sharded_pdws_pulsed_recording_and_phase_bug():
    """configs[4]-style pulsed recording through the channelizer, then the sharded extractor on 4 shards with
    the :114 column-1 phase bug reproduced (boundary pulses then need two columns)."""
    torch = _torch()
    M, P = 64, 12
    n = M * 40000
    iq, bw, fs = synth.pulsed_int16(n, M=M, seed=7)
    ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
    y = ch(iq, bw)
    ch.close()
    rows = y.shape[0]
    for bug in (False, True):
        kw = dict(fc=2.4e9, sampleStartTime=10.0, reproduce_phase_bug=bug)
        whole, nf = _pdws_on_matrix(y, fs, **kw)
        assert len(whole) > 3
        # put one boundary in the middle of the first pulse so at least one pulse straddles
        mid = int(whole[0].toa_row + whole[0].end_row) // 2
        bounds = sorted({0, mid, rows // 2, 3 * rows // 4, rows})
        for recs, nfs in _sharded_on_one_gpu(y, bounds, fs, **kw):
            _same_records(recs, whole)
            assert np.array_equal(nfs, nf)


# ---- event prediction (matlab/predict_event.m) ---------------------------------------
def _event_recording(seed, fs, n, t_peak, nf=0.004):
    """A dwell in which a pulsed emitter sweeps past: pulse amplitudes follow a parabola in dB that peaks
    at t_peak seconds into the file; one pulse reaches 0.95 of full scale so the script's gate (:52) opens."""
    rng = np.random.default_rng(seed)
    x = (rng.normal(0, nf, n) + 1j * rng.normal(0, nf, n)).astype(np.complex128)
    pri, pw = int(fs * 2e-3), int(fs * 60e-6)
    for a in range(pri // 2, n - pw, pri):
        tc = (a + pw / 2) / fs
        amp = 0.95 * 10.0 ** (-(40.0 * (tc - t_peak) ** 2) / 10.0)       # power-dB law, as the script's SNR (:95)
        x[a:a + pw] += amp * np.exp(2j * np.pi * 0.03 * np.arange(pw))
    return np.stack([np.clip(np.rint(x.real * 32768), -32768, 32767), np.clip(np.rint(x.imag * 32768), -32768, 32767)],
                    axis=1).astype(np.int16)


def test_predict_event_script_against_oracle(orc, tmp_path):
    """predict_event.m end to end on three dwells 4.5 s apart (plus one quiet dwell the :52 gate skips): PDWs with
    a single 20 dB threshold, parabola fit of SNR vs TOA, vertex = event time, next event from the median
    spacing -- against the oracle's PDWs + numpy.polyfit."""
    _torch()
    fs, n = 2e6, 200_000                                                  # 100 ms dwells
    starts = [1000.0, 1004.5, 1007.0, 1009.1]
    peaks = [0.047, 0.052, None, 0.044]
    paths, oevents = [], []
    for i, (t0, tp) in enumerate(zip(starts, peaks)):
        iq = _event_recording(40 + i, fs, n, tp) if tp is not None else (_event_recording(40 + i, fs, n, 0.05) // 8).astype(np.int16)
        p = str(tmp_path / f"dwell{i}.iq")
        pkg.write_iq(p, iq, fs=fs, fc=1.3e9, bitWidth=16, sampleStartTime=t0)
        paths.append(p)
        y = orc.unpack(iq, 16).reshape(-1, 1)
        if np.max(np.abs(y)) > 0.9:                                       # :52
            orecs, _ = orc.pdws(y, 1, snr_threshold_db=20.0, fc_hz=1.3e9, fs_sps=fs, t0=t0 - starts[0])
            oevents.append(orc.event_peak_time([r.toa_s for r in orecs], [r.snr_db for r in orecs])[0])
    res = pkg.predict_event(paths)
    assert res["pdws_per_file"][2] == 0 and all(c > 20 for i, c in enumerate(res["pdws_per_file"]) if i != 2)
    assert len(res["event"]) == len(oevents) == 3
    assert np.allclose(res["event"], oevents, atol=1e-6)
    for ev, t0, tp in zip(res["event"], [s for s, p in zip(starts, peaks) if p is not None], [p for p in peaks if p is not None]):
        assert abs(ev - (t0 - starts[0] + tp)) < 2e-3                     # the vertex sits where the emitter peaked
    assert abs(res["next_event"][0] - (res["event"][0] + 4.61962892466417)) < 1e-12
    assert abs(res["next_event"][-1] - orc.next_event_time(oevents)) < 1e-6


# ---- channelize_iq CLI: one recording, and the directory-watch mode ------------------
def _cli():
    import os
    p = os.path.join(os.path.dirname(pkg.LIB_PATH), "cli", "channelize_iq.out")
    assert os.path.exists(p), "build the CLI with __graft_entry__.build()"
    return p


def _expected(path, M, P):
    rec = pkg.read_iq(path)
    ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P), retain=True)
    y = ch(rec.iq, rec.bitWidth).copy()
    recs, _ = ch.pdws(rec.fs, rec.fc, rec.sampleStartTime)
    ch.close()
    return y, recs


def test_cli_single_recording_and_watch_mode(tmp_path):
    import os, subprocess, time
    _torch()
    M, P = 64, 12
    files = []
    for i in range(3):
        iq, bw, fs = synth.pulsed_int16(M * 4000, M=M, seed=300 + i)
        files.append((f"2024_01_0{i + 1}_00_00_00_000.iq", iq, bw, fs))     # Helper.cpp-style names: name order = time order
    # one recording
    one = str(tmp_path / files[0][0])
    pkg.write_iq(one, files[0][1], fs=files[0][3], fc=2.4e9, bitWidth=files[0][2], sampleStartTime=50.0)
    out = subprocess.run([_cli(), one, "0", str(P), "1", "15", str(tmp_path / "single")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    y, recs = _expected(one, M, P)                                          # channels = 0 -> fs*1e-6 = 64 (:31)
    got = np.fromfile(str(tmp_path / "single.cf32"), dtype=np.complex64).reshape(-1, M)
    assert np.array_equal(got.view(np.float32), y.view(np.float32))
    rows = open(str(tmp_path / "single.pdw.csv")).read().strip().splitlines()
    assert rows[0] == "toa_s,freq_hz,pw_s,snr_db,sat,amp,channel" and len(rows) - 1 == len(recs) > 0
    # watch mode: files dropped into the directory while the tool runs, the last one in two pieces
    watch, outd = tmp_path / "dwell", tmp_path / "out"
    watch.mkdir(); outd.mkdir()
    proc = subprocess.Popen([_cli(), str(watch), str(M), str(P), "1", "15", str(outd), "60"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        for name, iq, bw, fs in files[:2]:
            pkg.write_iq(str(watch / name), iq, fs=fs, fc=2.4e9, bitWidth=bw, sampleStartTime=50.0)
        name, iq, bw, fs = files[2]
        whole = str(tmp_path / "whole.iq")
        pkg.write_iq(whole, iq, fs=fs, fc=2.4e9, bitWidth=bw, sampleStartTime=50.0)
        data = open(whole, "rb").read()
        with open(str(watch / name), "wb") as f:                            # a recorder in the middle of its write
            f.write(data[:len(data) // 3]); f.flush()
            time.sleep(0.5)
            assert not (outd / (name[:-3] + ".pdw.csv")).exists()
            f.write(data[len(data) // 3:])
        deadline = time.time() + 120
        while time.time() < deadline and not all((outd / (n[:-3] + ".pdw.csv")).exists() for n, *_ in files):
            time.sleep(0.1)
        (watch / "stop").write_text("")
        so, se = proc.communicate(timeout=120)
    finally:
        if proc.poll() is None:
            proc.kill()
    assert proc.returncode == 0, se
    assert "Processed 3 recordings" in so
    for name, *_ in files:
        y, recs = _expected(str(watch / name), M, P)
        got = np.fromfile(str(outd / (name[:-3] + ".cf32")), dtype=np.complex64).reshape(-1, M)
        assert np.array_equal(got.view(np.float32), y.view(np.float32))
        assert len(open(str(outd / (name[:-3] + ".pdw.csv"))).read().strip().splitlines()) - 1 == len(recs)


def test_sharded_pdws_hysteresis_across_boundaries():
    """Wideband extractor (create_pdws.m: 18 dB up, 3 dB down) on shards: after a leading edge the samples that
    sit BETWEEN the two thresholds keep the pulse alive, so a shard made only of such samples hands its entry
    state on unchanged (exit code 2) and the FSM state has to be folded over several shards."""
    M, rows = 1, 4000
    rng = np.random.default_rng(11)
    y = (rng.normal(0, 0.001, rows) + 1j * rng.normal(0, 0.001, rows)).astype(np.complex64).reshape(-1, 1)
    def pulse(a, top, tail):                  # `top` strong rows, then a long plateau between the thresholds
        y[a:a + top, 0] += 0.5
        y[a + top:a + top + tail, 0] += 0.01  # ~ 8x the noise floor: above 3 dB, below 18 dB
    pulse(300, 40, 900)                       # plateau spans rows 340..1240
    pulse(2000, 10, 30)
    pulse(3100, 200, 500)
    kw = dict(SNR_THRESHOLD=18.0, TRAILING_EDGE_THRESHOLD=3.0)
    whole, nf = _pdws_on_matrix(y, 1e6, **kw)
    assert [(int(r.toa_row), int(r.end_row)) for r in whole][0][0] == 301 and whole[0].end_row > 1200 and len(whole) == 3
    for bounds in ([0, 500, 800, 1100, 4000], [0, 339, 340, 341, 2005, 4000], [0, 3150, 3400, 4000]):
        for recs, nfs in _sharded_on_one_gpu(y, bounds, 1e6, **kw):
            _same_records(recs, whole)
            assert np.array_equal(nfs, nf)


@pytest.mark.parametrize("M,seed", [(256, 100), (64, 102)])
def test_detector_on_the_oracles_thresholds(orc, M, seed):
    """the GPU takes its noise floor from fp32 |y| of fp32 channels, the oracle from doubles.  Feeding
    the GPU detector the ORACLE's floor (chz_pdw_shard_set_noise_floor) separates the two error sources: with the same
    thresholds on the same fp32 channel matrix every edge must fall on the same row (0 rows of tolerance) -- and with
    its own floor the GPU must reproduce the oracle's floor to fp32 accuracy and the same pulses."""
    torch = _torch()
    from sdr_channelizer_b200 import sharding
    n = M * 9000
    iq, bw, fs = synth.pulsed_int16(n, M=M, seed=seed)
    taps = pkg.design_prototype(M, 16)
    ch = pkg.Channelizer(M, taps=taps)
    ch.set_stream(torch.cuda.current_stream().cuda_stream)
    d_in = torch.from_numpy(iq).cuda()
    rows = n // M
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch.process_ptr(d_in.data_ptr(), n, bw, y.data_ptr(), rows)
    torch.cuda.synchronize()
    y_host = y.cpu().numpy()
    orecs, onf = orc.pdws(y_host.astype(np.complex128), M, fc_hz=1e9, fs_sps=fs, t0=2.0)   # oracle on the GPU's own fp32 channels
    assert len(orecs) >= 3
    sh = sharding.PdwShard(ch, y.data_ptr(), rows, 0, rows, fs, 1e9, 2.0)
    sh.set_noise_floor(onf)                                                               # oracle thresholds injected
    ev = sh.detect(np.zeros(M, np.uint8))
    pulses = sharding.pair_events(ev, M)
    assert [(p.channel_natural, p.toa_row, p.end_row) for p in pulses] == [(r.channel_natural, r.toa_row, r.end_row) for r in orecs]
    recs = [pkg.Pdw.from_buffer_copy(b) for b in sh.records(pulses)]
    _compare(recs, orecs, fs / M, tol_rows=0)
    own, nf = ch.pdws_ptr(y.data_ptr(), rows, fs, 1e9, 2.0)                               # the GPU's own median
    assert np.allclose(nf, onf, rtol=3e-7)
    _compare(own, orecs, fs / M, tol_rows=0)
    ch.close()
