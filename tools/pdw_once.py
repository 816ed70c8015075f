"""PDW extraction timing on synthetic pulsed files (profiling helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdr_channelizer_b200 as pkg
from tests import synth
M, P, fs = 256, 16, 56e6
n = 5_600_000 // M * M
for seed in (100, 101, 102, 103):
    iq, bw, _ = synth.pulsed_int16(n, M=M, seed=seed, fs=fs)
    d = torch.from_numpy(iq).cuda()
    rows = n // M
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
    if len(sys.argv) > 1:
        st = torch.cuda.Stream(); ch.set_stream(st.cuda_stream)
    ch.process_ptr(d.data_ptr(), n, bw, y.data_ptr(), rows); ch.synchronize()
    for i in range(2):
        t0 = time.perf_counter(); recs, nf = ch.pdws_ptr(y.data_ptr(), rows, fs); t1 = time.perf_counter()
        print("seed", seed, "pdws", len(recs), "ms", round((t1 - t0) * 1e3, 3), "max len", max((r.end_row - r.toa_row) for r in recs) if recs else 0)
    ch.close()
