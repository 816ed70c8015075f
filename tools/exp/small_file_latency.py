"""Where does the time of one small-file channelizer call go?  configs[4] geometry (M = 256, 5.6 M samples, int16).
Prints host time per call (call + synchronize), device time per call (CUDA events around the call), and the same
for back-to-back calls without a synchronize in between."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg

M, P = int(os.environ.get("M", "256")), int(os.environ.get("P", "16"))
n = int(os.environ.get("N", "5600000")) // M * M
rows = n // M
x = torch.randint(-2048, 2048, (n, 2), dtype=torch.int16, device="cuda")
y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
st = torch.cuda.current_stream(); ch.set_stream(st.cuda_stream)
for _ in range(3):
    ch.reset(); ch.process_ptr(x.data_ptr(), n, 12, y.data_ptr(), rows)
torch.cuda.synchronize()
K = 50
host = dev = 0.0
for _ in range(K):
    ch.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    ch.process_ptr(x.data_ptr(), n, 12, y.data_ptr(), rows)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    host += t1 - t0; dev += e0.elapsed_time(e1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
for _ in range(K):
    ch.reset(); ch.process_ptr(x.data_ptr(), n, 12, y.data_ptr(), rows)
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(json.dumps({"M": M, "P": P, "samples": n, "host_us_per_call_synced": round(host / K * 1e6, 1), "device_us_per_call_synced": round(dev / K * 1e3, 1),
                  "host_us_per_call_back_to_back": round((t1 - t0) / K * 1e6, 1), "device_us_per_call_back_to_back": round(e0.elapsed_time(e1) / K * 1e3, 1),
                  "roofline_us": round(n * 12 / 6456.2e9 * 1e6, 1)}))
