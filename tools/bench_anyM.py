"""Throughput of the functional any-M path (reference's natural M = fs*1e-6)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdr_channelizer_b200 as pkg
for M, P, bw in ((56, 12, 16), (560, 12, 16), (61, 12, 12), (64, 12, 12)):
    n = 56_000_000 // M * M
    x = torch.randint(-2000, 2000, (n, 2), dtype=torch.int16, device="cuda")
    rows = n // M
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(M, NumTapsPerBand=P)
    st = torch.cuda.Stream(); ch.set_stream(st.cuda_stream); torch.cuda.synchronize()
    with torch.cuda.stream(st):
        ch.process_ptr(x.data_ptr(), n, bw, y.data_ptr(), rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, bw, y.data_ptr(), rows)
        e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"M": M, "taps_per_band": P, "samples": n, "ms": ms, "MS_per_s": n / ms / 1e3}))
    ch.close()
