"""Deterministic synthetic recordings shaped like BASELINE.json's configs (the reference ships no
sample data and its generators are unseeded: matlab/generate_training_iq.m:13,18,22,24).

All generators return the RAW payload as the recorders write it: [N, 2] int8 / int16 (I, Q).
"""
import numpy as np


def _quantise(x, full_scale, lo, hi, dtype):
    re = np.clip(np.rint(x.real * full_scale), lo, hi)
    im = np.clip(np.rint(x.imag * full_scale), lo, hi)
    return np.stack([re, im], axis=1).astype(dtype)


def tones_complex(n, M, seed, centres=(1, 3, 6), amps=(0.5, 0.25, 0.1), off=(2.3, 0.3), sigma=0.01):
    """Tones at channel centres k*fs/M plus one off-centre tone and AWGN (configs[0] content)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    x = np.zeros(n, dtype=np.complex128)
    for k, a in zip(centres, amps):
        x += a * np.exp(2j * np.pi * (k % M) * t / M)
    x += off[1] * np.exp(2j * np.pi * off[0] * t / M)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x


def tones_int8(n, M=8, seed=1):
    """configs[0]: 8-bit file of tones.  -> (iq int8 [n,2], bitWidth 8)."""
    x = tones_complex(n, M, seed, amps=(0.35, 0.2, 0.1), off=(2.3, 0.2))
    return _quantise(x, 127.0, -128, 127, np.int8), 8


def tones_int16_q11(n, M=64, seed=2, ntones=8, sigma=0.05):
    """configs[1]: bladeRF-style 12-bit samples in int16 containers (SC16 Q11): 8 tones + AWGN,
    clipped to [-2048, 2047].  -> (iq int16 [n,2], bitWidth 12)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    x = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    ks = rng.choice(M, size=ntones, replace=False)
    for k in ks:
        f = (k + rng.uniform(-0.3, 0.3)) / M
        x += 0.08 * np.exp(2j * np.pi * (f * t + rng.uniform()))
    return _quantise(x, 2048.0, -2048, 2047, np.dtype("<i2")), 12


def noise_int16_full(n, seed=3):
    """configs[2]: full-scale int16 samples (b200mini 'sc16' wire format), uniform over the range."""
    rng = np.random.default_rng(seed)
    return rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16).astype(np.dtype("<i2")), 16


def pulsed_complex(n, fs, seed, sigma=0.005, amp=0.5):
    """One CW pulse train per file following matlab/generate_channelized_training_iq.m:12-68
    (f ~ U(-fs/2, fs/2), PW ~ U(10, 1000) us, PRI ~ U(max(10 us, PW), 10 ms), random start < PRI,
    phase restarting at each pulse), plus AWGN: the reference adds none, and without noise the median
    noise floor is 0 and the detector degenerates."""
    rng = np.random.default_rng(seed)
    f = -(fs / 2) + fs * rng.uniform()
    pw = 10e-6 + (1000e-6 - 10e-6) * rng.uniform()
    pri = max(10e-6, pw) + (10000e-6 - max(10e-6, pw)) * rng.uniform()
    npw, npri = int(round(fs * pw)), int(round(fs * pri))
    # keep several pulses inside short test files
    npri = min(npri, max(npw + 16, n // 6))
    npw = min(npw, max(8, npri // 2))
    start = int(rng.integers(1, max(2, npri)))
    x = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    ph = 2 * np.pi * f / fs * np.arange(npw)
    idx = start
    while idx < n:
        if idx + npw < n:
            x[idx:idx + npw] += amp * np.exp(1j * ph)
        idx += npri
    return x, dict(f=f, pw=npw / fs, pri=npri / fs, start=start)


def pulsed_int16(n, M=256, seed=100, fs=None):
    """configs[4] input: pulsed CW + AWGN quantised to int16 (bitWidth 16, generate_training_iq.m:95-98).
    fs defaults to M MHz so that M = fs*1e-6 as in create_pdws_channelized.m:31.
    -> (iq int16 [n,2], bitWidth 16, fs)."""
    fs = float(fs if fs is not None else M * 1e6)
    x, _ = pulsed_complex(n, fs, seed)
    return _quantise(x, 32768.0, -32768, 32767, np.dtype("<i2")), 16, fs


def rel_rms(a, b):
    """Relative RMS error of a against the reference b (north_star tolerance: <= 1e-5)."""
    den = np.mean(np.abs(b) ** 2)
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2) / den)) if den > 0 else float(np.max(np.abs(a - b)))
