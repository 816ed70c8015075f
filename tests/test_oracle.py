"""CPU tests of the oracle itself: pin it to what the reference's text and its one compilable artefact
(the IqPacket header struct) determine, and validate the channelizer restatement against independent
evaluations (direct form, scipy).  Channelizer parity with MATLAB's dsp.Channelizer is unpinned."""
import os
import struct

import numpy as np
import pytest
import scipy.signal as ss

from tests import synth


# ---- R1: header parser against the file written by the reference's own struct --------------------
def test_header_matches_reference_struct_12bit(orc, golden_dir):
    # tests/golden/ref_iqpacket_fmt3.iq was written by oracle/_ref/ref_iqpacket_writer, which fills the
    # reference's IqPacket (cpp/IqPacket.h:9-25) and dumps it like blade_record_iq_12bit.cpp:320-323.
    info, iq = orc.read_iq(os.path.join(golden_dir, "ref_iqpacket_fmt3.iq"))
    assert (info.magic, info.format, info.header_bytes) == (0x03030303, 3, 112)
    assert info.link_speed == 5000 and info.fc_hz == 5_800_000_000 and info.bw_hz == 56_000_000
    assert info.fs_sps == 61_440_000 and info.gain_db == 37.5 and info.num_samples == 37
    assert info.bit_width == 12 and info.spare0 == 0 and info.bytes_per_sample == 4
    assert info.board_name == b"bladerf2" and info.serial_number == b"0123456789abcdef"
    assert info.fpga_version == b"0.15.0" and info.fw_version == b"2.4.0"
    assert info.sample_start_time == 1700000000.123456
    i = np.arange(37)
    assert np.array_equal(iq[:, 0], (i * 113 - 2048).astype(np.int16))
    assert np.array_equal(iq[:, 1], (2047 - i * 97).astype(np.int16))


def test_header_matches_reference_struct_8bit(orc, golden_dir):
    info, iq = orc.read_iq(os.path.join(golden_dir, "ref_iqpacket_fmt3_8bit.iq"))
    assert info.bit_width == 8 and info.bytes_per_sample == 2 and iq.dtype == np.int8
    i = np.arange(37)
    assert np.array_equal(iq[:, 0], (i * 7 - 128).astype(np.int8))
    assert np.array_equal(iq[:, 1], (127 - i * 5).astype(np.int8))


def _fmt1_bytes(n=5, bit_width=16, fs=56_000_000):
    # exactly what matlab/generate_training_iq.m:107-125 writes (format 1, 104-byte header)
    hdr = struct.pack("<IIIIIIII", 0x01010101, 1, 0, fs, fs, 0, n, bit_width)
    hdr += b"simulated" + bytes(64 - 9) + struct.pack("<d", 12.5)
    payload = np.arange(2 * n, dtype="<i2").tobytes()
    return hdr + payload


def test_header_format1_and_2(orc):
    rc, info = orc.parse_header(_fmt1_bytes())
    assert rc == 0 and info.format == 1 and info.header_bytes == 104 and info.fs_sps == 56_000_000
    assert info.board_name == b"simulated" and info.sample_start_time == 12.5 and info.num_samples == 5
    # format 2 (bladeRF magic 0x02020202, blade_record_iq_12bit.cpp:248): u64 fc, u32 gain, spare0
    hdr = struct.pack("<IIQIIIIII", 0x02020202, 3, 2_400_000_000, 20_000_000, 61_440_000, 30, 3, 12, 0)
    hdr += bytes(64) + struct.pack("<d", 1.0)
    rc, info = orc.parse_header(hdr + bytes(12))
    assert rc == 0 and info.format == 2 and info.gain_db == 30.0 and info.fc_hz == 2_400_000_000
    # magic 0 ("big endian") is read as format 2 (convert_my_iq_to_mat.m:43-45)
    rc, info = orc.parse_header(struct.pack("<I", 0) + hdr[4:] + bytes(12))
    assert rc == 0 and info.format == 2


def test_header_errors(orc):
    good = _fmt1_bytes()
    assert orc.parse_header(b"\x05\x05\x05\x05" + good[4:])[0] == -3          # Unsupported endianness (:55-56)
    bad_bw = bytearray(good); bad_bw[28:32] = struct.pack("<I", 17)
    assert orc.parse_header(bytes(bad_bw))[0] == -4                            # Unsupported bit width (:96-97)
    bad_bw[28:32] = struct.pack("<I", 0)
    assert orc.parse_header(bytes(bad_bw))[0] == -4
    assert orc.parse_header(good[:-4])[0] == -5                                # assert(length(iq)==numSamples) (:102)
    assert orc.parse_header(good + bytes(4))[0] == -5
    assert orc.parse_header(good + bytes(3))[0] == 0                           # fread [2,inf] drops a ragged tail


# ---- R2: normalise -----------------------------------------------------------------------------------
def test_unpack_all_values(orc):
    v8 = np.arange(-128, 128, dtype=np.int8)
    iq = np.stack([v8, v8[::-1]], axis=1)
    x = orc.unpack(iq, 8)
    assert np.array_equal(x.real, v8 / 128.0) and np.array_equal(x.imag, v8[::-1] / 128.0)
    v = np.array([-32768, -2048, -1, 0, 1, 2047, 32767], dtype=np.int16)
    iq = np.stack([v, v], axis=1)
    assert np.array_equal(orc.unpack(iq, 12).real, v / 2048.0)
    assert np.array_equal(orc.unpack(iq, 16).real, v / 32768.0)


# ---- R4: default prototype -----------------------------------------------------------------------------
@pytest.mark.parametrize("M,P", [(8, 8), (64, 12), (64, 16), (256, 12)])
def test_prototype_vs_scipy_firwin(orc, M, P):
    h = orc.design_prototype(M, P, 80.0)
    beta = 0.1102 * (80.0 - 8.7)
    ref = ss.firwin(M * P, 1.0 / M, window=("kaiser", beta), scale=False)
    ref /= ref.sum()
    assert abs(h.sum() - 1.0) < 1e-12 and np.max(np.abs(h - ref)) < 1e-12
    # stop band (beyond the adjacent channel centre) is >= 75 dB down
    H = np.abs(np.fft.rfft(h, 64 * M * P))
    f = np.arange(len(H)) / (64 * M * P)
    assert 20 * np.log10(H[f >= 1.0 / M].max()) < -75.0


# ---- R5: channelizer restatement ---------------------------------------------------------------------
@pytest.mark.parametrize("M,P,os_", [(8, 8, 1), (8, 4, 2), (16, 3, 2), (64, 16, 1), (56, 12, 1), (32, 12, 2)])
def test_polyphase_equals_direct_form(orc, M, P, os_):
    rng = np.random.default_rng(M * 100 + P)
    h = orc.design_prototype(M, P)
    x = rng.standard_normal(M * 37 + 5) + 1j * rng.standard_normal(M * 37 + 5)
    a, b = orc.channelize(x, M, h, os_), orc.channelize_direct(x, M, h, os_)
    assert a.shape == (len(x) // (M // os_), M)
    assert np.max(np.abs(a - b)) < 1e-12


def test_channel0_is_lfilter_then_downsample(orc):
    M, P = 16, 12
    rng = np.random.default_rng(5)
    h = orc.design_prototype(M, P)
    x = rng.standard_normal(M * 50) + 1j * rng.standard_normal(M * 50)
    y = orc.channelize(x, M, h)
    assert np.max(np.abs(y[:, 0] - ss.lfilter(h, 1.0, x)[::M])) < 1e-13


def test_matches_numpy_fft_of_branch_sums(orc):
    # y[m] = M * ifft(u[m]) with u_p[m] = sum_q h[qM+p] x[mM - qM - p]
    M, P = 32, 8
    rng = np.random.default_rng(9)
    h = orc.design_prototype(M, P)
    x = rng.standard_normal(M * 40) + 1j * rng.standard_normal(M * 40)
    y = orc.channelize(x, M, h)
    xp = np.concatenate([np.zeros(M * P, complex), x])
    for m in (0, 3, 17, 39):
        u = np.array([sum(h[q * M + p] * xp[M * P + m * M - q * M - p] for q in range(P)) for p in range(M)])
        assert np.max(np.abs(y[m] - M * np.fft.ifft(u))) < 1e-12


@pytest.mark.parametrize("os_", [1, 2])
def test_on_bin_tone_lands_in_one_channel_with_unit_gain(orc, os_):
    M, P, k0 = 64, 16, 5
    h = orc.design_prototype(M, P)
    n = np.arange(M * 200)
    y = orc.channelize(np.exp(2j * np.pi * k0 * n / M), M, h, os_)
    last = y[-1]
    assert abs(abs(last[k0]) - 1.0) < 1e-9
    assert np.max(np.abs(np.delete(last, k0))) < 1e-4
    # the output is basebanded: a tone at the channel centre has constant phase row to row (also 2x oversampled)
    assert np.max(np.abs(np.angle(y[-20:, k0] / y[-21:-1, k0]))) < 1e-9


def test_impulse_response_rows_are_polyphase_taps(orc):
    M, P = 8, 8
    h = orc.design_prototype(M, P)
    x = np.zeros(M * 20, complex); x[0] = 1.0
    y = orc.channelize(x, M, h)
    for m in range(P):    # u_p[m] = h[mM + p] only for p = 0 (x[0] reaches branch 0 at row q = m)
        assert np.max(np.abs(y[m] - h[m * M])) < 1e-15
    assert np.max(np.abs(y[P:])) == 0.0


def test_row_range_and_raw_entry(orc):
    iq, bw = synth.tones_int16_q11(64 * 300, 64, seed=2)
    h = orc.design_prototype(64, 16)
    full = orc.channelize(orc.unpack(iq, bw), 64, h)
    part = orc.channelize_raw(iq, bw, 64, h, row0=100, nrows=50)
    assert np.max(np.abs(part - full[100:150])) < 1e-15


def test_center_frequencies(orc):
    f = orc.center_frequencies(8, 8e6)
    assert np.array_equal(f, (np.arange(8) - 4) * 1e6)


# ---- R9-R11: PDW state machine and medians -----------------------------------------------------------
def test_fsm_equality_toggles(orc):
    # create_pdws_channelized.m:88 uses >=, :94 uses <= on the same threshold: equality flips the state
    assert orc.fsm_trace([0, 0, 2, 2, 2, 0, 3, 0], 2) == [(3, 4), (5, 6), (7, 8)]


def test_fsm_open_pulse_at_end_is_dropped(orc):
    assert orc.fsm_trace([0, 5, 5, 0, 5, 5], 2) == [(2, 4)]
    assert orc.fsm_trace([5, 5, 5], 2) == []
    assert orc.fsm_trace([], 2) == []


def test_median_even_odd(orc):
    assert orc.median([3, 1, 2]) == 2 and orc.median([4, 1, 3, 2]) == 2.5 and orc.median([7]) == 7


def _pulse_matrix(M=8, rows=400):
    """Hand-built channel matrix: unit-ish noise plus two pulses with known rows in known channels."""
    rng = np.random.default_rng(11)
    y = 0.01 * (rng.standard_normal((rows, M)) + 1j * rng.standard_normal((rows, M)))
    k1, k2 = 2, 6
    n1 = np.arange(100, 151)
    y[100:151, k1] += 0.8 * np.exp(1j * (0.3 * (n1 - 100)))          # +0.3 rad/row
    n2 = np.arange(200, 231)
    y[200:231, k2] += 0.5 * np.exp(-1j * (0.2 * (n2 - 200)))         # -0.2 rad/row
    y[210, k2] = 1.0 + 0.0j                                           # saturated sample inside the pulse
    return y, k1, k2


def test_pdws_known_answer(orc):
    M, fs, fc, t0 = 8, 8e6, 1e9, 1000.0
    y, k1, k2 = _pulse_matrix(M)
    recs, nf = orc.pdws(y, M, fc_hz=fc, fs_sps=fs, t0=t0)
    assert len(recs) == 2
    fs_dec = fs / M
    # output order: shifted channel ascending (:79); natural k=6 is shifted column 2, k=2 is column 6
    a, b = recs
    assert (a.channel_natural, a.channel) == (k2, (k2 + M // 2) % M) and (b.channel_natural, b.channel) == (k1, (k1 + 4) % 8)
    # leading edge at 0-based row 100 -> 1-based 101; trailing edge = first row back at/below threshold
    assert (b.toa_row, b.end_row) == (101, 152) and (a.toa_row, a.end_row) == (201, 232)
    assert b.toa_s == 101 / fs_dec + t0 and b.pw_s == (152 - 101) / fs_dec            # :98, :110
    assert abs(b.amp - 0.8) < 0.03 and abs(b.snr_db - 10 * np.log10(b.amp / nf[k1])) < 1e-12
    f_expect = fc + orc.center_frequencies(M, fs)[b.channel] + fs_dec * np.degrees(0.3) / 360.0
    assert abs(b.freq_hz - f_expect) < 0.02 * fs_dec
    assert a.saturated == 1 and b.saturated == 0
    assert np.allclose(nf, np.median(np.abs(y), axis=0))


def test_pdws_saturation_ignored_on_edges(orc):
    M = 8
    y, k1, k2 = _pulse_matrix(M)
    y[210, k2] = 0.5                    # remove the interior saturation
    y[200, k2] = 1.0                    # leading-edge sample: not checked (:88-92)
    recs, _ = orc.pdws(y, M, fs_sps=8e6)
    assert [r.saturated for r in recs] == [0, 0]


def test_pdws_phase_bug_flag(orc):
    # :114 indexes phase(toa:jj) with one subscript => always shifted column 1 (natural channel M/2)
    M = 8
    y, k1, k2 = _pulse_matrix(M)
    y[:, M // 2] = 0.01 * np.exp(1j * 0.5 * np.arange(y.shape[0]))    # steady +0.5 rad/row in natural channel M/2
    good, _ = orc.pdws(y, M, fs_sps=8e6)
    bug, _ = orc.pdws(y, M, fs_sps=8e6, reproduce_phase_bug=True)
    fs_dec = 1e6
    cf = orc.center_frequencies(M, 8e6)
    for r in bug:
        assert abs((r.freq_hz - cf[r.channel]) - fs_dec * np.degrees(0.5) / 360.0) < 1.0
    assert abs((good[1].freq_hz - cf[good[1].channel]) - fs_dec * np.degrees(0.3) / 360.0) < 0.02 * fs_dec


def test_wideband_hysteresis_thresholds(orc):
    """matlab/create_pdws.m:45-47,58,63: leading edge at 18 dB over the median, trailing edge at 3 dB.
    One column (the un-channelized stream), D = 1."""
    n = 2000
    x = np.full(n, 0.01 + 0j)
    x[500:600] = 0.9                      # strong pulse
    x[600:650] = 0.05                     # sags below the leading but above the 3 dB trailing threshold: still active
    x[650] = 0.0101                       # <= trailing threshold (0.01 * 10^0.3 = 0.01995) -> trailing edge
    x[900:905] = 0.8
    recs, nf = orc.pdws(x.reshape(-1, 1), 1, snr_threshold_db=18.0, fs_sps=1e6, trailing_snr_threshold_db=3.0)
    assert nf[0] == 0.01
    assert [(r.toa_row, r.end_row) for r in recs] == [(501, 651), (901, 906)]
    # with the single 18 dB threshold the first pulse ends as soon as it sags
    recs1, _ = orc.pdws(x.reshape(-1, 1), 1, snr_threshold_db=18.0, fs_sps=1e6)
    assert [(r.toa_row, r.end_row) for r in recs1] == [(501, 601), (901, 906)]
