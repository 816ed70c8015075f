"""PDW stage (K4) at headline scale: configs[1] geometry output (rows x 64 channels) with pulsed tones.
Prints per-stage wall-clock (CHZ_PDW_TRACE) and the algorithmic re-read traffic rate."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CHZ_PDW_TRACE"] = "1"
import torch
import sdr_channelizer_b200 as pkg

M, P = 64, 16
n = int(float(sys.argv[1]) * 61_440_000) // M * M if len(sys.argv) > 1 else 61_440_000 * 2
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(3)
x = (torch.randn((n, 2), device=dev, generator=g) * 20).to(torch.int16)          # noise, sigma ~ 0.01 FS (12-bit)
t = torch.arange(n, device=dev, dtype=torch.float64)
for k, period, width in ((5, 200_000, 40_000), (20, 333_333, 10_000), (41, 1_000_000, 300_000)):
    on = ((t % period) < width)
    ph = 2 * torch.pi * torch.frac(t * ((k + 0.13) / M))
    x[:, 0] += (600 * torch.cos(ph) * on).to(torch.int16)
    x[:, 1] += (600 * torch.sin(ph) * on).to(torch.int16)
del t
rows = n // M
y = torch.empty((rows, M), dtype=torch.complex64, device=dev)
ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
ch.process_ptr(x.data_ptr(), n, 12, y.data_ptr(), rows); ch.synchronize()
for i in range(3):
    t0 = time.perf_counter(); recs, nf = ch.pdws_ptr(y.data_ptr(), rows, 61.44e6); dt = time.perf_counter() - t0
    print(f"run {i}: {len(recs)} PDWs in {dt * 1e3:.2f} ms; y = {rows * M * 8 / 1e9:.2f} GB -> {4 * rows * M * 8 / dt / 1e9:.0f} GB/s over 4 passes; "
          f"{n / dt / 1e6:.0f} input MS/s", flush=True)
