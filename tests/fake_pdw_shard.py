"""CPU stand-in for sdr_channelizer_b200.sharding.PdwShard (test infrastructure only): the same stage
semantics in numpy, so that the distributed orchestration of create_pdws_sharded -- histogram sums, FSM
state folding, boundary-pulse stitching, record merging -- can run under a world_size > 1 gloo group on a
machine without a GPU.  The product never imports this."""
import numpy as np
import torch

from sdr_channelizer_b200 import _lib

KBINS = 2048
SHIFT = (20, 9, 0)
BMASK = (0x7FF, 0x7FF, 0x1FF)
PMASK = (0x0, 0xFFF00000, 0xFFFFFE00)


def _mag32(y):
    y = np.asarray(y, dtype=np.complex64)
    return np.sqrt(y.real.astype(np.float32) ** 2 + y.imag.astype(np.float32) ** 2, dtype=np.float32)


class FakePdwShard:
    def __init__(self, y, row_offset, total_rows, D, fs, fc=0.0, t0=0.0, snr_db=15.0, sat_level=0.9999):
        self.y = np.ascontiguousarray(y, dtype=np.complex64)
        self.nrows, self.M = self.y.shape
        self.row_offset, self.total_rows, self.D = int(row_offset), int(total_rows), int(D)
        self.fs, self.fc, self.t0, self.snr_db, self.sat = fs, fc, t0, snr_db, sat_level
        self.bug = False
        self.bits = _mag32(self.y).view(np.uint32)
        self.prefix = np.zeros((self.M, 2), dtype=np.uint32)
        self.rank = np.zeros((self.M, 2), dtype=np.int64)
        self._hist = None

    # -- median ---------------------------------------------------------------------------------
    def hist(self, p):
        h = np.zeros((self.M, 2, KBINS), dtype=np.int32)
        for k in range(self.M):
            b = self.bits[:, k]
            for slot in range(2):
                if slot == 1 and (p == 0 or self.prefix[k, 0] == self.prefix[k, 1]):
                    continue
                sel = b[(b & PMASK[p]) == self.prefix[k, slot]] if p else b
                np.add.at(h[k, slot], (sel >> SHIFT[p]) & BMASK[p], 1)
        self._hist = torch.from_numpy(h.reshape(-1))
        return self._hist

    def select(self, p):
        h = self._hist.numpy().reshape(self.M, 2, KBINS)
        if p == 0:
            self.rank[:, 0] = (self.total_rows - 1) // 2
            self.rank[:, 1] = self.total_rows // 2
        for k in range(self.M):
            split = p > 0 and self.prefix[k, 0] != self.prefix[k, 1]
            new = []
            for slot in range(2):
                row = h[k, 1 if (slot == 1 and split) else 0]
                c = np.cumsum(row)
                b = int(np.searchsorted(c, self.rank[k, slot], side="right"))
                new.append((b, self.rank[k, slot] - (c[b - 1] if b else 0)))
            for slot, (b, r) in enumerate(new):
                self.prefix[k, slot] |= np.uint32(b << SHIFT[p])
                self.rank[k, slot] = r

    def thresholds(self):
        lo = self.prefix[:, 0].copy().view(np.float32).astype(np.float64)
        hi = self.prefix[:, 1].copy().view(np.float32).astype(np.float64)
        self.nf = 0.5 * (lo + hi)
        self.thr = self.nf * 10.0 ** (self.snr_db / 10.0)

    def noise_floor(self):
        return self.nf.copy()

    # -- edges -----------------------------------------------------------------------------------
    def _mag(self):
        return _mag32(self.y).astype(np.float64)

    def exit_state(self):
        mag = self._mag()
        code = np.zeros(self.M, dtype=np.uint8)
        for k in range(self.M):
            flips, c = False, 2
            for j in range(self.nrows - 1, -1, -1):
                m = mag[j, k]
                if m == self.thr[k]:
                    flips = not flips
                    continue
                c = 1 if m > self.thr[k] else 0
                break
            code[k] = (3 if flips else 2) if c == 2 else ((1 - c) if flips else c)
        return code

    def detect(self, entry):
        mag = self._mag()
        ev = []
        for k in range(self.M):
            active = bool(entry[k])
            chs = (k + self.M // 2) % self.M
            for j in range(self.nrows):
                m = mag[j, k]
                if not active and m >= self.thr[k]:
                    active = True
                    ev.append((chs << 40) | ((j + 1 + self.row_offset) << 1))
                elif active and m <= self.thr[k]:
                    active = False
                    ev.append((chs << 40) | ((j + 1 + self.row_offset) << 1) | 1)
        return np.asarray(ev, dtype=np.uint64)

    # -- records ----------------------------------------------------------------------------------
    def _records(self, mat, row_offset, pulses):
        out = []
        fs_dec = self.fs / self.D
        for p in pulses:
            a, b = p.toa_row - 1 - row_offset, p.end_row - 1 - row_offset
            col = mat[a:b + 1, p.col]
            ph = np.degrees(np.angle(mat[a:b + 1, p.col_phase].astype(np.complex128)))
            d = np.diff(ph)
            d = np.where(d < -180, d + 360, d)
            d = np.where(d > 180, d - 360, d)
            inner = mat[a + 1:b, p.col]
            r = _lib.Pdw()
            k = p.channel_natural
            c = (k + self.M // 2) % self.M
            r.amp = float(np.median(_mag32(col).astype(np.float64)))
            r.noise_floor = float(self.nf[k])
            r.snr_db = 10 * np.log10(r.amp / self.nf[k])
            r.toa_s = p.toa_row / fs_dec + self.t0
            r.pw_s = (p.end_row - p.toa_row) / fs_dec
            med = float(np.median(d.astype(np.float32).astype(np.float64)))
            r.freq_hz = (self.fc + (c - self.M // 2) * self.fs / self.M) + (fs_dec / (360.0 / med) if med else 0.0)
            r.channel, r.channel_natural = c, k
            r.toa_row, r.end_row = p.toa_row, p.end_row
            r.saturated = int(np.any((np.abs(inner.real) >= self.sat) | (np.abs(inner.imag) >= self.sat)))
            out.append(bytes(r))
        return out

    def records(self, pulses):
        return self._records(self.y, self.row_offset, pulses)

    def column_segment(self, cols, row_lo, row_hi):
        return self.y[row_lo - 1 - self.row_offset:row_hi - self.row_offset][:, list(cols)].copy()

    def records_from_matrix(self, mat, row_offset, pulses):
        return self._records(np.asarray(mat, dtype=np.complex64), row_offset, pulses)
