"""Per-stage GPU times of the PDW extractor on a LARGE channel matrix (default 3.84 M rows x 64 channels = 1.97 GB of
noise with a few pulses): run with CHZ_PDW_TRACE=1; CHZ_PDW_HIST_NARROW=1 selects the 4-channel histogram kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg
M = int(os.environ.get("M", "64")); rows = int(os.environ.get("ROWS", "3840000"))
torch.manual_seed(1)
y = torch.randn((rows, M, 2), device="cuda") * 0.01
y[1000:1400, 3] += 0.5; y[500000:500900, 17] += 0.4; y[rows - 5000:rows - 4000, M - 1] += 0.6
y = torch.view_as_complex(y.contiguous())
ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, 16))
ch.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    t0 = time.perf_counter(); recs, nf = ch.pdws_ptr(y.data_ptr(), rows, 61.44e6); dt = time.perf_counter() - t0
print("records", len(recs), "ms", round(dt * 1e3, 3), "nf0", float(nf[0]))
