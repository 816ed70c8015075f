"""numpy model of k_chan_ring's index arithmetic (chz_ring.cuh): frames, columns, delta, branch-0 fix-up, tile
positions and the in-place DIF passes, checked against the oracle.  CPU only; a design aid, not a test."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle as orc

M, NT, R = 1024, 512, 8


def tpad(pos):
    return pos + ((pos >> 7) << 1)


def model(x, taps, os_, P, nrows):
    D = M // os_
    TS = M + 2 * (M // 128)
    h = taps.reshape(P, M)
    tw = np.exp(2j * np.pi * np.arange(M) / M)

    def frame(a):          # A_a = x[aM, aM+M), zeros outside
        out = np.zeros(M, dtype=np.complex128)
        lo = a * M
        for c in range(M):
            if 0 <= lo + c < len(x):
                out[c] = x[lo + c]
        return out

    y = np.zeros((nrows, M), dtype=np.complex128)
    nsteps = ((nrows - 1) // os_ + 1 + R - 1) // R
    t = np.arange(NT)
    b1, b2 = 2 * t + 1, (2 * t + 2) & (M - 1)
    for k in range(nsteps):
        a0 = k * R
        ring = np.stack([frame(a0 - 16 + i) for i in range(24)])     # rows i = 0..23
        for ph in range(os_):
            tile = np.zeros((R, TS), dtype=np.complex128)
            low = (ph == 1) & (t < D // 2)
            cl = np.where(ph == 1, np.where(low, D - 2 * t - 2, M + D - 2 * t - 2), M - 2 * t - 2)
            delta = low.astype(int)
            acc1 = np.zeros((R, NT), dtype=np.complex128)
            acc2 = np.zeros((R, NT), dtype=np.complex128)
            for ii in range(P + R - 1):
                j = ii + 16 - P
                xa = ring[j + delta, cl]
                xb = ring[j + delta, cl + 1]
                for r in range(R):
                    q = r + P - 1 - ii
                    if 0 <= q < P:
                        acc2[r] += h[q, b2] * xa
                        acc1[r] += h[q, b1] * xb
            shift = D if ph else 0
            for r in range(R):
                tile[r, tpad((b1 - shift) & (M - 1))] = acc1[r]
                tile[r, tpad((b2 - shift) & (M - 1))] = acc2[r]
            for r in range(R):       # branch 0 fix-up
                acc = 0
                for q in range(P - 1, -1, -1):
                    i = 16 + r - q
                    acc += h[q, 0] * ring[i, D if ph else 0]
                tile[r, tpad((0 - shift) & (M - 1))] = acc
            # pass 0
            for tt in range(NT):
                j, rr = tt & 127, tt >> 7
                for hh in range(2):
                    base = j
                    row = tile[rr + 4 * hh]
                    v = np.array([row[base + q * 130] for q in range(8)])
                    v = np.fft.ifft(v) * 8
                    for q in range(1, 8):
                        v[q] *= tw[(j * q) & (M - 1)]
                    for q in range(8):
                        row[base + q * 130] = v[q]
            # pass 1
            for tt in range(NT):
                j1, kb, rr = tt & 15, (tt >> 4) & 7, tt >> 7
                for hh in range(2):
                    row = tile[rr + 4 * hh]
                    base = kb * 130 + j1
                    v = np.array([row[base + q * 16] for q in range(8)])
                    v = np.fft.ifft(v) * 8
                    for q in range(1, 8):
                        v[q] *= tw[((j1 * q) << 3) & (M - 1)]
                    for q in range(8):
                        row[base + q * 16] = v[q]
            # pass 2
            for tt in range(NT):
                row_i, b = tt >> 6, tt & 63
                kb, kc = b & 7, b >> 3
                base = kb * 130 + kc * 16
                v = np.fft.ifft(tile[row_i, base:base + 16]) * 16
                m = (a0 + row_i) * os_ + ph
                if m < nrows:
                    for q in range(16):
                        y[m, b + q * 64] = v[q]
    return y


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for os_, P in ((1, 16), (2, 16), (2, 12), (1, 8)):
        n = M * 21 + 517
        x = (rng.integers(-2000, 2000, n) + 1j * rng.integers(-2000, 2000, n)).astype(np.complex128)
        taps = rng.uniform(-1, 1, M * P)
        ref = orc.channelize(x, M, taps, os_)
        got = model(x, taps, os_, P, ref.shape[0])
        err = np.sqrt(np.mean(np.abs(got - ref) ** 2) / np.mean(np.abs(ref) ** 2))
        print(f"os={os_} P={P} rows={ref.shape[0]} rel_rms={err:.3e}")
