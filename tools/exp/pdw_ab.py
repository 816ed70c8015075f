import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import sdr_channelizer_b200 as pkg
from tests import synth
M = 1
n = 8 * 9000 * 8
iq, bw, fs = synth.pulsed_int16(n, M=8, seed=105)
taps = np.ones(1, np.float32)
res = []
for ep in (0, 1):
    ch = pkg.Channelizer(M, taps=taps, retain=True)
    ch.set_option(pkg.CHZ_OPT_PDW_EVENT_PATH, ep)
    ch(iq, bw)
    recs, nf = ch.pdws(fs, 2.4e9, 17.0, SNR_THRESHOLD=18.0, TRAILING_EDGE_THRESHOLD=3.0)
    res.append([(r.toa_row, r.end_row, r.amp, r.freq_hz, r.saturated) for r in recs])
    ch.close()
print(len(res[0]), len(res[1]))
for a, b in zip(res[0], res[1]):
    if a != b: print("DIFF", a, b)
print(res[1][:5])
