"""CPU tests of the C-ABI library: it loads, exports every symbol include/channelizer.h declares, and its
host-side pieces (reader, writer, prototype designer, error paths) agree with the oracle and the golden
files.  No compute entry point is exercised here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

import sdr_channelizer_b200 as pkg
from sdr_channelizer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "channelizer.h")).read()
    declared = set(re.findall(r"\b(chz_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = pkg.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert L.chz_abi_version() == 1


def test_strerror_and_error_codes():
    L = pkg.lib()
    assert L.chz_strerror(0) == b"ok"
    assert b"endianness" in L.chz_strerror(_lib.CHZ_EFORMAT)
    assert b"bit width" in L.chz_strerror(_lib.CHZ_EBITWIDTH)
    assert b"no CPU path" in L.chz_strerror(_lib.CHZ_ENODEVICE)


def test_reader_on_reference_struct_golden(golden_dir, orc):
    for name in ("ref_iqpacket_fmt3.iq", "ref_iqpacket_fmt3_8bit.iq"):
        path = os.path.join(golden_dir, name)
        rec = pkg.read_iq(path)
        oinfo, oiq = orc.read_iq(path)
        assert np.array_equal(rec.iq, oiq) and rec.iq.dtype == oiq.dtype
        for f, _ in _lib.IqInfo._fields_:
            assert getattr(rec.info, f) == getattr(oinfo, f), f
    rec = pkg.read_iq(os.path.join(golden_dir, "ref_iqpacket_fmt3.iq"))
    # variable names of convert_my_iq_to_mat.m:118
    assert (rec.fs, rec.fc, rec.bw, rec.gain, rec.bitWidth) == (61.44e6, 5.8e9, 56e6, 37.5, 12)
    assert rec.sampleStartTime == 1700000000.123456 and rec.boardName == "bladerf2" and rec.serialNo == "0123456789abcdef"
    assert rec.dur == 37 / 61.44e6 and rec.fileFormat == 3


@pytest.mark.parametrize("fmt", [1, 2, 3])
@pytest.mark.parametrize("bw", [8, 12, 16])
def test_write_read_roundtrip_all_formats(tmp_path, orc, fmt, bw):
    rng = np.random.default_rng(fmt * 10 + bw)
    dt = np.int8 if bw <= 8 else np.int16
    iq = rng.integers(-100, 100, size=(33, 2)).astype(dt)
    path = str(tmp_path / "r.iq")
    pkg.write_iq(path, iq, fs=56_000_000, fc=915_000_000, bw=40_000_000, gain=30, bitWidth=bw,
                 sampleStartTime=123.25, fileFormat=fmt, linkSpeed=5000, boardName="b200mini", serialNo="abc")
    assert os.path.getsize(path) == (104 if fmt == 1 else 112) + iq.nbytes
    rec = pkg.read_iq(path)
    oinfo, oiq = orc.read_iq(path)          # the oracle's parser agrees byte for byte
    assert np.array_equal(rec.iq, iq) and np.array_equal(oiq, iq)
    assert rec.fileFormat == fmt == oinfo.format and rec.bitWidth == bw and rec.fs == 56e6 and rec.fc == 915e6
    assert rec.gain == 30.0 and rec.sampleStartTime == 123.25 and rec.boardName == "b200mini"
    # byte offsets of cpp/IqPacket.h (format >= 2): fsSps@20, numSamples@28, bitWidth@32, start@104
    raw = open(path, "rb").read()
    if fmt >= 2:
        assert struct.unpack_from("<I", raw, 20)[0] == 56_000_000 and struct.unpack_from("<I", raw, 28)[0] == 33
        assert struct.unpack_from("<I", raw, 32)[0] == bw and struct.unpack_from("<d", raw, 104)[0] == 123.25
        assert struct.unpack_from("<Q", raw, 8)[0] == 915_000_000
    else:
        assert struct.unpack_from("<I", raw, 8)[0] == 915_000_000 and struct.unpack_from("<d", raw, 96)[0] == 123.25


def test_reader_errors(tmp_path):
    good = struct.pack("<IIQIIfIII", 0x03030303, 0, 1, 2, 3, 1.0, 4, 16, 0) + bytes(64) + struct.pack("<d", 0.0) + bytes(16)

    def rc_of(data):
        p = str(tmp_path / "x.iq")
        open(p, "wb").write(data)
        return pkg.lib().chz_open_iq(p.encode(), None, C.byref(_lib.IqInfo()))

    assert rc_of(good) == 0
    assert rc_of(b"\x09\x09\x09\x09" + good[4:]) == _lib.CHZ_EFORMAT
    bad = bytearray(good); bad[32:36] = struct.pack("<I", 17)
    assert rc_of(bytes(bad)) == _lib.CHZ_EBITWIDTH
    assert rc_of(good[:-4]) == _lib.CHZ_ESIZE and rc_of(good + bytes(4)) == _lib.CHZ_ESIZE
    assert rc_of(good[:50]) == _lib.CHZ_EIO
    assert pkg.lib().chz_open_iq(b"/nonexistent/file.iq", None, C.byref(_lib.IqInfo())) == _lib.CHZ_EIO
    with pytest.raises(pkg.ChannelizerError, match="endianness"):
        open(str(tmp_path / "y.iq"), "wb").write(b"\x09\x09\x09\x09" + good[4:])
        pkg.read_iq(str(tmp_path / "y.iq"))


@pytest.mark.parametrize("M,P", [(8, 8), (64, 12), (64, 16), (1024, 16)])
def test_prototype_matches_oracle(orc, M, P):
    h = pkg.design_prototype(M, P, 80.0)
    ref = orc.design_prototype(M, P, 80.0)
    assert h.dtype == np.float32 and np.array_equal(h, ref.astype(np.float32))


def test_create_argument_checks_and_no_cpu_fallback():
    L = pkg.lib()
    h = C.c_void_p()
    for M in (0, 4097, 8192):
        assert L.chz_create(M, None, 0, 1, C.byref(h)) == _lib.CHZ_EINVAL
    assert L.chz_create(7, None, 0, 2, C.byref(h)) == _lib.CHZ_EINVAL      # 2x oversampling needs an even M
    assert L.chz_create(64, None, 0, 3, C.byref(h)) == _lib.CHZ_EINVAL
    taps = np.zeros(100, dtype=np.float32)
    assert L.chz_create(64, taps.ctypes.data_as(C.c_void_p), 100, 1, C.byref(h)) == _lib.CHZ_EINVAL
    if not _has_gpu():
        # the product must fail loudly, not fall back to a CPU implementation
        assert L.chz_create(64, None, 0, 1, C.byref(h)) == _lib.CHZ_ENODEVICE
        with pytest.raises(pkg.ChannelizerError, match="no CPU path"):
            pkg.Channelizer(64)
        buf = np.zeros(8, dtype=np.int16)
        assert L.chz_unpack_dev(buf.ctypes.data_as(C.c_void_p), 4, 16, buf.ctypes.data_as(C.c_void_p), None) == _lib.CHZ_ENODEVICE


def test_product_never_imports_the_oracle():
    # only tests/, __graft_entry__.smoke() and bench.py may touch oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sdr_channelizer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f == "__init__.py" and False, os.path.join(dirpath, f)


# ---- host-only entry points: event prediction (matlab/predict_event.m:125-138) -------------------------
def test_event_peak_time_against_numpy_polyfit(orc):
    import sdr_channelizer_b200 as pkg
    rng = np.random.default_rng(3)
    t = np.sort(rng.uniform(0.0, 9.0, 40))
    v = 22.0 - 0.8 * (t - 4.25) ** 2 + rng.normal(0, 0.2, t.size)
    tp, vp, coef = pkg.event_peak_time(t, v)
    otp, ovp = orc.event_peak_time(t, v)
    assert abs(tp - otp) < 1e-9 and abs(vp - ovp) < 1e-9
    assert np.allclose(coef[::-1], np.polyfit(t, v, 2), rtol=1e-9, atol=1e-9)
    # an exact parabola is recovered exactly (to rounding)
    tp, vp, _ = pkg.event_peak_time(t, 5.0 - 2.0 * (t - 3.0) ** 2)
    assert abs(tp - 3.0) < 1e-10 and abs(vp - 5.0) < 1e-9
    for bad_t, bad_v in (([1.0, 2.0], [1.0, 2.0]), ([1.0, 1.0, 1.0, 1.0], [1.0, 2.0, 3.0, 4.0]), ([0.0, 1.0, 2.0, 3.0], [0.0, 1.0, 2.0, 3.0])):
        with pytest.raises(pkg.ChannelizerError):      # too few points / rank deficient / no curvature
            pkg.event_peak_time(bad_t, bad_v)


def test_next_event_time_median_rules(orc):
    import sdr_channelizer_b200 as pkg
    ev = [10.0, 14.5, 19.2, 23.7, 28.6]                 # differences 4.5 4.7 4.5 4.9 -> median 4.6
    assert abs(pkg.next_event_time(ev) - orc.next_event_time(ev)) < 1e-12
    assert abs(pkg.next_event_time(ev) - (28.6 + 4.6)) < 1e-12
    assert abs(pkg.next_event_time(ev, upper_median=True) - (28.6 + 4.7)) < 1e-12     # usrp_predict_event.cpp:366
    assert abs(pkg.next_event_time(ev[:4]) - (23.7 + 4.5)) < 1e-12                     # odd count: the middle one
    assert pkg.next_event_time([7.0]) == 7.0 + 4.61962892466417                        # predict_event.m:137
