"""Parity tests proper: the CUDA path through the C ABI against the CPU oracle, on a real B200.
Tolerances are north_star's: unpack bit-exact; channel outputs rel. RMS <= 1e-5 vs the double oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import sdr_channelizer_b200 as pkg
from sdr_channelizer_b200 import _lib
from tests import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5   # relative RMS, BASELINE.json north_star


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch


def _force_path(ch, path):
    """CHZ_OPT_FORCE_PATH; the round-1 experiment kernels (paths 3-10) exist only in `make EXPERIMENTS=1` builds."""
    try:
        ch.set_option(_lib.CHZ_OPT_FORCE_PATH, path)
    except pkg.ChannelizerError as e:
        if e.code == _lib.CHZ_EINVAL and 3 <= path <= 10:
            pytest.skip("experiment kernels not built (make -C sdr_channelizer_b200/csrc EXPERIMENTS=1)")
        raise


def _gen(kind, n, M, seed):
    if kind == "i8":
        return synth.tones_int8(n, M, seed)
    if kind == "q11":
        return synth.tones_int16_q11(n, M, seed)
    return synth.noise_int16_full(n, seed)


# ---- K1 ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bw", [8, 5, 12, 16])
def test_unpack_bit_exact_all_values(orc, bw):
    torch = _torch()
    if bw <= 8:
        v = np.arange(-128, 128, dtype=np.int8)
        iq = np.stack([np.tile(v, 3), np.tile(v[::-1], 3)], axis=1)
    else:
        v = np.arange(-32768, 32768, dtype=np.int16)
        iq = np.stack([v, np.roll(v[::-1], 77)], axis=1)
    d_in = torch.from_numpy(iq).cuda()
    d_out = torch.empty((iq.shape[0], 2), dtype=torch.float32, device="cuda")
    pkg.unpack_ptr(d_in.data_ptr(), iq.shape[0], bw, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    ref = orc.unpack(iq, bw)
    assert np.array_equal(got[:, 0].astype(np.float64), ref.real) and np.array_equal(got[:, 1].astype(np.float64), ref.imag)


@pytest.mark.parametrize("bw,n,off_in,off_out", [(12, 1001, 0, 0), (12, 1000, 1, 0), (12, 999, 0, 1), (8, 777, 1, 1),
                                                 (8, 4096, 0, 0), (16, 1, 0, 0), (16, 2, 3, 0)])
def test_unpack_odd_counts_and_misaligned_buffers(orc, bw, n, off_in, off_out):
    """K1 takes two samples per thread (8-byte loads, 16-byte stores) when the buffers allow it: an odd count leaves
    one sample to the scalar kernel, a misaligned input or output sends the whole call there."""
    torch = _torch()
    rng = np.random.default_rng(n + bw)
    lim = 2 ** (bw - 1)
    iq = rng.integers(-lim, lim, size=(n + off_in, 2)).astype(np.int8 if bw <= 8 else np.int16)
    d_in = torch.from_numpy(iq).cuda()
    d_out = torch.full((n + off_out + 1, 2), -7.0, dtype=torch.float32, device="cuda")
    pkg.unpack_ptr(d_in[off_in:].data_ptr(), n, bw, d_out[off_out:].data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    ref = orc.unpack(iq[off_in:], bw)
    g = got[off_out:off_out + n]
    assert np.array_equal(g[:, 0].astype(np.float64), ref.real) and np.array_equal(g[:, 1].astype(np.float64), ref.imag)
    assert np.all(got[:off_out] == -7.0) and np.all(got[off_out + n:] == -7.0)       # nothing written outside


# ---- K3 against cuFFT ----------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 56, 560,   # 56 = 8*7, 560 = 16*5*7: radix-7/5 passes
                               2, 6, 7, 12, 40, 48, 100, 112, 120, 200, 243, 768, 1000, 3000, 3584,   # run-time mixed-radix plans
                               22, 61])                                                         # O(M^2) DFT (prime factor > 7)
def test_fft_stage_matches_cufft(M):
    torch = _torch()
    rows = 37 if M >= 1024 else 333
    g = torch.Generator(device="cuda").manual_seed(M)
    u = torch.randn((rows, M), dtype=torch.complex64, device="cuda", generator=g)
    y = torch.empty_like(u)
    ch = pkg.Channelizer(M, NumTapsPerBand=8)
    ch.set_stream(torch.cuda.current_stream().cuda_stream)
    ch.fft_rows_ptr(u.data_ptr(), y.data_ptr(), rows)
    torch.cuda.synchronize()
    ref = torch.fft.ifft(u.to(torch.complex128), dim=1) * M          # cuFFT, e^{+j} exponent, unnormalised
    err = (torch.linalg.norm(y.to(torch.complex128) - ref) / torch.linalg.norm(ref)).item()
    assert err < 2e-6, err
    ch.close()


# ---- fused / split channelizer against the oracle --------------------------------------------------------
CASES = [
    # M, P, oversample, input kind, samples
    (8, 8, 1, "i8", 1_000_000),          # configs[0]
    (8, 12, 2, "i8", 100_003),
    (16, 12, 1, "q11", 50_000),
    (32, 16, 2, "q11", 64_000),
    (64, 16, 1, "q11", 64 * 5000 + 17),  # configs[1] geometry
    (64, 12, 1, "i8", 64 * 3000),
    (64, 8, 2, "full", 64 * 3000),
    (128, 16, 1, "q11", 128 * 1500),
    (256, 16, 1, "full", 256 * 1100),    # configs[4] geometry
    (256, 12, 2, "q11", 256 * 600),
    (512, 16, 1, "q11", 512 * 400),
    (1024, 16, 2, "full", 1024 * 300),   # configs[2] geometry
    (4096, 16, 1, "q11", 4096 * 100),    # configs[3] geometry
    (2048, 12, 2, "i8", 2048 * 64),
    (1024, 12, 1, "i8", 1024 * 211 + 5),   # ring kernel, 8-bit recording, toolbox-default 12 taps per band
    (1024, 8, 2, "q11", 1024 * 150),
    (1024, 16, 1, "full", 1024 * 1300),  # ring kernel on many CTAs (warm-up frames of every CTA's run)
    (64, 4, 1, "q11", 64 * 2000),        # split path (no fused instantiation for P=4)
    (64, 5, 1, "q11", 64 * 500),         # generic-P fallback
]


@pytest.mark.parametrize("M,P,os_,kind,n", CASES)
def test_channelizer_matches_oracle(orc, M, P, os_, kind, n):
    _torch()
    iq, bw = _gen(kind, n, M, seed=M + P)
    taps = pkg.design_prototype(M, P)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    y = ch(iq, bw)
    ref = orc.channelize_raw(iq, bw, M, taps.astype(np.float64), os_)
    assert y.shape == ref.shape == (n // (M // os_), M)
    err = synth.rel_rms(y, ref)
    assert err <= TOL, err
    # row 0 depends on x[0] only: exact product of one tap and one sample
    assert abs(y[0, 0] - ref[0, 0]) <= 1e-6 * abs(ref[0, 0]) + 1e-12
    ch.close()


@pytest.mark.parametrize("M,P,os_,path", [(64, 16, 1, 1), (64, 16, 2, 1), (64, 12, 1, 1), (8, 8, 1, 1), (256, 16, 1, 1),
                                          (512, 8, 2, 1), (32, 12, 2, 1), (128, 16, 1, 1), (16, 16, 1, 1),
                                          (64, 16, 1, 2), (1024, 16, 2, 0), (4096, 8, 1, 0), (64, 24, 1, 0), (64, 32, 2, 0),
                                          (4096, 16, 1, 3), (2048, 16, 2, 3), (1024, 8, 1, 3), (64, 16, 1, 4), (64, 16, 2, 4), (64, 12, 1, 4), (64, 8, 2, 4),
                                          (1024, 16, 1, 5), (1024, 16, 2, 5), (1024, 12, 2, 5), (1024, 8, 1, 5),
                                          (4096, 16, 1, 6), (2048, 16, 2, 6), (1024, 8, 1, 6), (1024, 12, 2, 6), (1024, 16, 2, 6), (2048, 12, 1, 6),
                                          (1024, 16, 2, 7), (2048, 16, 1, 7), (4096, 16, 1, 7), (1024, 16, 1, 7),
                                          (1024, 16, 2, 10), (2048, 16, 1, 10), (4096, 16, 1, 10), (1024, 16, 1, 10), (2048, 16, 2, 10), (4096, 16, 2, 10),
                                          (1024, 16, 2, 8), (2048, 16, 1, 8), (4096, 16, 1, 8), (1024, 16, 1, 8), (4096, 16, 2, 9), (1024, 8, 1, 9), (2048, 16, 1, 9),
                                          (1024, 16, 1, 11), (1024, 16, 2, 11), (1024, 12, 1, 11), (1024, 12, 2, 11), (1024, 8, 1, 11), (1024, 8, 2, 11), (1024, 16, 2, 2),
                                          (56, 12, 1, 1), (56, 16, 2, 1), (56, 8, 1, 1), (56, 12, 2, 2), (560, 12, 1, 1), (560, 16, 2, 1), (560, 8, 1, 1), (560, 12, 2, 2)])
def test_random_taps_every_tap_index_matters(orc, M, P, os_, path):
    """A designed prototype has tiny end taps, which would hide a mis-indexed tap or window slot
    below the 1e-5 tolerance; with random taps of equal weight any such slip is an O(1/P) error."""
    _torch()
    rng = np.random.default_rng(M * 7 + P)
    taps = (rng.uniform(0.5, 1.0, M * P) * rng.choice([-1.0, 1.0], M * P) / P).astype(np.float32)
    n = M * (5 * P + 3) * 4 + 11
    iq, bw = synth.noise_int16_full(n, seed=P)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    _force_path(ch, path)
    y = ch(iq, bw)
    ref = orc.channelize_raw(iq, bw, M, taps.astype(np.float64), os_)
    assert y.shape == ref.shape
    assert synth.rel_rms(y, ref) <= 2e-6
    assert np.max(np.abs(y - ref)) <= 1e-5 * np.max(np.abs(ref))
    ch.close()


@pytest.mark.parametrize("M,P,os_", [(64, 16, 1), (64, 16, 2), (256, 12, 1), (8, 8, 1), (56, 12, 1), (56, 12, 2), (560, 12, 1)])
def test_split_path_equals_fused_path(M, P, os_):
    _torch()
    iq, bw = synth.tones_int16_q11(M * 900 + 5, M, seed=3)
    taps = pkg.design_prototype(M, P)
    outs = []
    for path in (1, 2):
        ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
        _force_path(ch, path)
        outs.append(ch(iq, bw))
        ch.close()
    assert synth.rel_rms(outs[0], outs[1]) < 1e-6


@pytest.mark.parametrize("M,P,os_", [(1024, 16, 2), (1024, 16, 1), (2048, 12, 2), (4096, 16, 1), (4096, 8, 2)])
def test_pipelined_path_is_bit_identical_to_split_path(M, P, os_):
    """Path 6 (one persistent launch, FIR and in-place FFT tasks from one ticket queue) runs the split
    path's arithmetic, so its rows must match bit for bit -- one shot and cut into ragged calls (odd row
    counts move the pair alignment, so FFT tasks then depend on two span groups)."""
    _torch()
    n = M * 1500 // os_ + 13
    iq, bw = synth.tones_int16_q11(n, M, seed=21)
    taps = pkg.design_prototype(M, P)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    _force_path(ch, 2)
    whole = ch(iq, bw).copy()
    ch.close()
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    _force_path(ch, 6)
    one = ch(iq, bw).copy()
    assert np.array_equal(one.view(np.float32), whole.view(np.float32))
    ch.reset()
    D = M // os_
    parts, pos = [], 0
    for step in [D * 201 + 7, D * 3, 1, D * 333 + D // 2, D * 70, n]:
        parts.append(ch(iq[pos:pos + step], bw).copy())
        pos += step
        if pos >= n:
            break
    got = np.concatenate(parts, axis=0)
    assert got.shape == whole.shape and np.array_equal(got.view(np.float32), whole.view(np.float32))
    ch.close()


@pytest.mark.parametrize("M,P,os_", [(64, 16, 1), (32, 12, 2), (1024, 16, 2), (8, 8, 1)])
def test_streaming_chunks_are_bit_identical_to_one_shot(M, P, os_):
    """Stateful like the System object (channelizer_example.m:50-56): ragged chunk sizes, including
    ones shorter than a frame, give exactly the rows of a single call."""
    _torch()
    n = M * 700 + 13
    iq, bw = synth.tones_int16_q11(n, M, seed=9)
    taps = pkg.design_prototype(M, P)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    whole = ch(iq, bw).copy()
    ch.reset()
    rng = np.random.default_rng(1)
    parts, pos = [], 0
    while pos < n:
        step = int(rng.choice([1, 3, M // 2, M, M + 1, 5 * M + 7, 100 * M]))
        parts.append(ch(iq[pos:pos + step], bw).copy())
        pos += step
    got = np.concatenate(parts, axis=0)
    assert got.shape == whole.shape and np.array_equal(got.view(np.float32), whole.view(np.float32))
    ch.close()


@pytest.mark.parametrize("M,P,os_,world", [(64, 16, 1, 2), (64, 16, 1, 8), (1024, 16, 2, 4), (4096, 16, 1, 8), (256, 12, 2, 3)])
def test_time_shards_bit_identical_to_unsharded(M, P, os_, world):
    """configs[3] scaling path on one GPU: each shard (halo of taps-1 samples fed first, its rows
    discarded) reproduces exactly the rows of the unsharded run, so stitching is a plain concatenation."""
    _torch()
    n = M * 150 * world + 3 * M + 1
    iq, bw = synth.tones_int16_q11(n, M, seed=13)
    taps = pkg.design_prototype(M, P)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    whole = ch(iq, bw).copy()
    parts = []
    for sh in pkg.plan_time_shards(n, M, M * P, os_, world):
        ch.reset()
        parts.append(ch(iq[sh.sample_begin:sh.sample_end], bw)[sh.discard_rows:].copy())
        assert parts[-1].shape[0] == sh.rows
    got = pkg.stitch_rows(parts)
    assert got.shape == whole.shape and np.array_equal(got.view(np.float32), whole.view(np.float32))
    ch.close()


def test_host_chunked_pipeline_equals_single_chunk():
    _torch()
    M = 64
    iq, bw = synth.tones_int16_q11(M * 5000, M, seed=4)
    taps = pkg.design_prototype(M, 16)
    ch = pkg.Channelizer(M, taps=taps)
    a = ch(iq, bw).copy()
    ch.reset()
    ch.set_option(_lib.CHZ_OPT_CHUNK_ROWS, 333)      # many small pipelined chunks
    ch.set_option(_lib.CHZ_OPT_RETAIN, 0)
    b = ch(iq, bw).copy()
    assert np.array_equal(a.view(np.float32), b.view(np.float32))
    ch.close()


def test_impulse_and_tone_known_answers():
    _torch()
    M, P = 64, 16
    taps = pkg.design_prototype(M, P)
    ch = pkg.Channelizer(M, taps=taps)
    iq = np.zeros((M * 40, 2), dtype=np.int16); iq[0, 0] = 2047
    y = ch(iq, 12)
    for m in range(P):      # impulse -> row m is the constant taps[m*M] * x[0] in every channel
        assert np.allclose(y[m], taps[m * M] * (2047 / 2048), rtol=1e-6, atol=1e-12)
    assert np.all(y[P:] == 0)
    ch.reset()
    k0 = 11
    n = np.arange(M * 400)
    x = 0.5 * np.exp(2j * np.pi * k0 * n / M)
    iq = np.stack([np.rint(x.real * 32768), np.rint(x.imag * 32768)], axis=1).astype(np.int16)
    y = ch(iq, 16)
    assert abs(abs(y[-1, k0]) - 0.5) < 1e-4 and np.max(np.abs(np.delete(y[-1], k0))) < 1e-3
    ch.close()


@pytest.mark.parametrize("M,P,os_,bw", [(64, 16, 1, 12), (8, 8, 2, 8), (1024, 16, 2, 16), (64, 32, 1, 10), (256, 12, 1, 5)])
def test_edge_lengths_ragged_and_tiny(orc, M, P, os_, bw):
    """Empty, shorter than a frame, exactly one frame, one sample more, shorter than the prototype,
    ragged tails (the reference trims to a multiple of M, create_pdws_channelized.m:52-54)."""
    _torch()
    D, L = M // os_, M * P
    rng = np.random.default_rng(M + bw)
    taps = pkg.design_prototype(M, P)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    lim = 2 ** (bw - 1)
    dt = np.int8 if bw <= 8 else np.int16
    for n in (0, 1, D - 1, D, D + 1, M, 2 * M + 3, L - 1, L, L + D + 1, 3 * L + 5):
        iq = rng.integers(-lim, lim, size=(n, 2)).astype(dt)
        ch.reset()
        y = ch(iq, bw)
        ref = orc.channelize_raw(iq, bw, M, taps.astype(np.float64), os_)
        assert y.shape == ref.shape == (n // D, M), n
        if y.size:
            assert synth.rel_rms(y, ref) <= TOL, n
    ch.close()


@pytest.mark.parametrize("M,P,os_,kind", [(56, 12, 1, "q11"), (56, 12, 2, "full"), (12, 8, 1, "i8"), (7, 5, 1, "full"),
                                         (100, 16, 2, "full"), (560, 12, 1, "q11"), (4, 8, 1, "i8"), (2, 3, 2, "full"),
                                         (40, 12, 1, "q11"), (48, 16, 2, "full"), (80, 12, 1, "i8"), (120, 8, 1, "q11"), (200, 12, 2, "full"),
                                         (768, 12, 1, "q11"), (1000, 12, 1, "full"), (61, 12, 1, "q11"), (22, 12, 2, "full"), (250, 5, 1, "q11")])
def test_non_power_of_two_channel_counts(orc, M, P, os_, kind):
    """The reference's own M is fs*1e-6 (56 for the b200mini at 56 MS/s, create_pdws_channelized.m:31);
    56 and 560 have compile-time radix-7/5 plans and run the fused kernel; other channel counts with prime
    factors <= 7 (40, 48, 100, 200 ... as natural as 56 for other sample rates) run the register-window FIR +
    a run-time mixed-radix row FFT; the rest (61, 22) the direct FIR / O(M^2) DFT.  Same tolerance, and
    streaming stays bit-identical."""
    _torch()
    n = M * 300 + 5
    iq, bw = _gen(kind, n, M, seed=M)
    taps = pkg.design_prototype(M, P)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    y = ch(iq, bw).copy()
    ref = orc.channelize_raw(iq, bw, M, taps.astype(np.float64), os_)
    assert y.shape == ref.shape == (n // (M // os_), M)
    assert synth.rel_rms(y, ref) <= TOL
    ch.reset()
    parts = [ch(iq[a:b], bw).copy() for a, b in ((0, 3), (3, M * 7 + 1), (M * 7 + 1, M * 200), (M * 200, n))]
    assert np.array_equal(np.concatenate(parts).view(np.float32), y.view(np.float32))
    ch.close()


def test_capacity_and_state_errors():
    torch = _torch()
    ch = pkg.Channelizer(64, NumTapsPerBand=16)
    iq = np.zeros((64 * 10, 2), dtype=np.int16)
    out = np.empty((3, 64), dtype=np.complex64)
    n = C.c_uint64(0)
    rc = pkg.lib().chz_process(ch.handle, iq.ctypes.data_as(C.c_void_p), 640, 12, out.ctypes.data_as(C.c_void_p), 3, C.byref(n))
    assert rc == _lib.CHZ_ECAPACITY and n.value == 10
    assert ch(iq, 12).shape == (10, 64)
    with pytest.raises(pkg.ChannelizerError):      # bit width may not change mid-stream
        ch(iq, 16)
    ch.reset()
    assert ch(iq, 16).shape == (10, 64)
    assert ch(np.zeros((0, 2), dtype=np.int16), 16).shape == (0, 64)
    with pytest.raises(pkg.ChannelizerError):
        ch(iq, 17)
    ch.close()


# ---- size-independent properties at (near) BASELINE size ------------------------------------------------
@pytest.mark.parametrize("M,P,os_,bw,nframes", [(64, 16, 1, 12, 589_000),      # configs[1] geometry, 37.7 M samples
                                               (4096, 16, 1, 12, 22_000),     # configs[3] geometry, 90.1 M samples
                                               (1024, 16, 2, 16, 40_000),     # configs[2] geometry, 41.0 M samples
                                               (256, 16, 1, 16, 175_000)])    # configs[4] geometry, 44.8 M samples
def test_large_run_spot_rows_and_linearity(orc, M, P, os_, bw, nframes):
    """BASELINE geometries at sizes the oracle cannot run whole: random rows are checked against the
    oracle evaluated on just the samples those rows touch, and the transform is linear:
    chan(a) + chan(b) == chan(a + b) when a + b does not clip."""
    torch = _torch()
    n = M * nframes
    D, L = M // os_, M * P
    lim = 2 ** (bw - 1) // 2 - 1
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(-lim, lim, (n, 2), dtype=torch.int16, device="cuda", generator=g)
    b = torch.randint(-lim, lim, (n, 2), dtype=torch.int16, device="cuda", generator=g)
    taps = pkg.design_prototype(M, P)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=os_)
    ch.set_stream(torch.cuda.current_stream().cuda_stream)
    rows = n // D
    outs = []
    for src in (a, b, a + b):
        ch.reset()
        y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
        assert ch.process_ptr(src.data_ptr(), n, bw, y.data_ptr(), rows) == rows
        outs.append(y)
    torch.cuda.synchronize()
    lin = (torch.linalg.norm(outs[0] + outs[1] - outs[2]) / torch.linalg.norm(outs[2])).item()
    assert lin < 1e-6, lin
    rng = np.random.default_rng(0)
    for m in [0, 1, 2, os_ * P - 1, rows - 2, rows - 1] + [int(v) for v in rng.integers(P, rows, 18)]:
        lo = max(0, m * D - (L - 1))
        seg = a[lo:m * D + 1].cpu().numpy()                                   # x[mD-L+1 .. mD], clipped at 0
        seg_full = np.concatenate([np.zeros((L - len(seg), 2), np.int16), seg])
        # window in which x[mD] sits at index idx = L + (mD mod M): same frame and rotation phase as in
        # the full run, complete FIR history before it; its row idx/D is row m of the full run
        idx = L + (m * D) % M
        win = np.concatenate([np.zeros((idx - (L - 1), 2), np.int16), seg_full, np.zeros((D - 1, 2), np.int16)])
        ref = orc.channelize_raw(win, bw, M, taps.astype(np.float64), os_)[idx // D]
        got = outs[0][m].cpu().numpy()
        assert synth.rel_rms(got, ref) <= TOL, m
    ch.close()


def test_stft_of_spectrogram_script(orc, tmp_path):
    """matlab/spectrogram_my_iq.m:104-115: 768-point Hamming STFT without overlap == the M = 768, one-tap-per-band
    filterbank (run-time mixed-radix plan 16*16*3) up to a fixed phase factor; against numpy's FFT of the
    windowed segments of the oracle's normalised samples."""
    _torch()
    fs, n = 56e6, 768 * 300 + 401
    iq, bw = synth.tones_int16_q11(n, 64, seed=17)
    s, f, t = pkg.stft(iq, bw, fs)
    x = orc.unpack(iq, bw)
    nseg = n // 768
    ref = np.fft.fftshift(np.fft.fft(x[:nseg * 768].reshape(nseg, 768) * np.hamming(768), axis=1), axes=1).T
    assert s.shape == ref.shape == (768, nseg)
    assert synth.rel_rms(s, ref) <= TOL
    assert np.allclose(f, (np.arange(768) - 384) * fs / 768) and np.allclose(t, (np.arange(nseg) * 768 + 384) / fs)
    path = str(tmp_path / "spec.iq")
    pkg.write_iq(path, iq, fs=fs, fc=1.0e9, bitWidth=bw)
    sp = pkg.spectrogram_my_iq(path)
    assert sp["power"].shape == (768, nseg) and np.allclose(sp["power"], np.abs(ref) ** 2, rtol=1e-4, atol=1e-9 * np.max(np.abs(ref)) ** 2)
    assert sp["f_hz"][384] == 1.0e9


# ---- R7 centerFrequencies and the channelizer_example.m script ---------------------------------------------
@pytest.mark.parametrize("M", [8, 56, 7, 1, 64, 255])
def test_center_frequencies_product(orc, M):
    """centerFrequencies(channelizer, fs) as the scripts use it, against the shifted columns
    (create_pdws_channelized.m:42,60,80; channelizer_example.m:60): chz_channel_freq per natural channel and the
    Python mirror's centerFrequencies against the oracle, even and odd M."""
    _torch()
    fs = 56e6
    taps = np.ones(M, np.float32) / M
    ch = pkg.Channelizer(M, taps=taps)
    ref = orc.center_frequencies(M, fs)                              # ascending, shifted column order
    got = ch.centerFrequencies(fs)
    assert got.shape == (M,) and np.allclose(got, ref, rtol=0, atol=1e-6)
    assert np.all(np.diff(got) > 0) or M == 1
    nat = np.array([pkg.lib().chz_channel_freq(ch.handle, k, fs) for k in range(M)])
    for k in range(M):                                               # natural channel k sits in shifted column (k + M//2) % M (:60)
        assert abs(nat[k] - ref[(k + M // 2) % M]) <= 1e-6
    assert nat[0] == 0.0 and np.isnan(pkg.lib().chz_channel_freq(ch.handle, M, fs))
    ch.close()


def test_channelizer_example_script(orc, tmp_path):
    """matlab/channelizer_example.m:18-61: conjugated input (:23), ONE stateful channelizer fed overlapping 5 ms windows
    stepped by 100 frames (:50-56), abs, fftshift (:58), axes (:60-61) -- against the oracle run on the conjugated
    samples with the FIR history each call inherits from the window before it."""
    _torch()
    fs, fc, M, P = 8e6, 915e6, 8, 12
    n = 64_000
    x = synth.tones_complex(n, M, seed=5, centres=(1, 3, 6), amps=(0.4, 0.2, 0.1), off=(2.3, 0.15))
    iq = np.stack([np.clip(np.rint(x.real * 2047), -2048, 2047), np.clip(np.rint(x.imag * 2047), -2048, 2047)], axis=1).astype(np.int16)
    path = str(tmp_path / "demo.iq")
    pkg.write_iq(path, iq, fs=fs, fc=fc, bitWidth=12, sampleStartTime=5.0)
    taps = pkg.design_prototype(M, P)
    frames = list(pkg.channelizer_example(path, taps=taps))
    samples, step, L = int(5e-3 * fs), 100 * M, M * P
    starts = [ii for ii in range(1, n + 1, step) if ii + samples - 1 <= n]
    assert len(frames) == len(starts) >= 20
    xc = np.conj(orc.unpack(iq, 12))                                 # :18-23
    h = taps.astype(np.float64)
    prev = np.zeros(0, dtype=np.complex128)
    for (f, t, z), ii in zip(frames, starts):
        win = xc[ii - 1:ii - 1 + samples]
        tail = prev[-(L // M + 1) * M:]                              # whole frames covering the L-1 samples of FIR history
        ref = np.abs(orc.channelize(np.concatenate([tail, win]), M, h))[len(tail) // M:]
        ref = np.fft.fftshift(ref, axes=1)                           # :58
        assert z.shape == ref.shape == (samples // M, M)
        assert synth.rel_rms(z, ref) <= TOL
        assert np.allclose(f, (fc - orc.center_frequencies(M, fs)) * 1e-6)             # :60
        assert np.allclose(t, ii / fs + np.arange(samples // M) * M / fs)              # :61
        prev = win
    # a tone at +3 fs/M in the recording shows up at -3 fs/M after the conjugate, i.e. at f = fc + 3 MHz on the script's axis
    z = frames[-1][2]
    col = int(np.argmax(z.mean(axis=0)))
    assert abs(frames[-1][0][col] - (fc * 1e-6 + 1.0)) < 1e-9        # strongest tone: centre 1 -> -1 MHz offset -> f = fc + 1


def test_ring_kernel_random_shapes_are_repeatable_and_match_the_split_path():
    """tools/exp/ring_stress.py: random lengths, chunkings, bit widths, oversampling, taps per band and input
    alignments through the M = 1024 kernel (FIR warps and FFT warps handing tiles over through named barriers, TMA
    ring with mbarriers).  Three runs of a case must be bit-identical -- compute-sanitizer's racecheck is closed on
    this pool, and a race would show as nondeterminism -- and agree with the split path to rel-RMS 1e-5."""
    import subprocess
    import sys
    _torch()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "exp", "ring_stress.py"), "24"], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, SEED="11"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
