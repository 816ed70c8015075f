"""Fused kernel across channel counts on one B200 (560 M int16 samples, critically sampled)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for M in [int(m) for m in os.environ.get("CHZ_BENCH_MS", "8,16,32,56,64,128,256,512,560").split(",")]:
    n = 560_000_000 // M * M
    x = torch.randint(-2000, 2000, (n, 2), dtype=torch.int16, device="cuda")
    rows = n // M
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
    st = torch.cuda.Stream(); ch.set_stream(st.cuda_stream); torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, 16, y.data_ptr(), rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, 16, y.data_ptr(), rows)
        e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"M": M, "P": P, "ms": round(ms, 4), "GS_per_s": round(n / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(12 * n / (ms * 1e-3) / 6456.2e9, 4)}), flush=True)
    ch.close(); del x, y; torch.cuda.empty_cache()
