"""Throughput at the reference's natural channel counts (M = fs*1e-6 = 56, and 560 for 0.1 MHz bins): the
radix-7/5 plans (fused kernel), the split path (CHZ_BENCH_PATH=2) and the functional any-M path (M = 61)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdr_channelizer_b200 as pkg
for M, P, bw in ((56, 12, 16), (560, 12, 16), (61, 12, 12), (64, 12, 12), (40, 12, 16), (100, 12, 16), (200, 12, 16), (768, 12, 16)):
    n = 560_000_000 // M * M if M != 61 else 56_000_000 // M * M     # 10 s at 56 MS/s
    x = torch.randint(-2000, 2000, (n, 2), dtype=torch.int16, device="cuda")
    rows = n // M
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(M, NumTapsPerBand=P)
    if os.environ.get("CHZ_BENCH_PATH"):
        ch.set_option(pkg.CHZ_OPT_FORCE_PATH, int(os.environ["CHZ_BENCH_PATH"]))
    st = torch.cuda.Stream(); ch.set_stream(st.cuda_stream); torch.cuda.synchronize()
    with torch.cuda.stream(st):
        ch.process_ptr(x.data_ptr(), n, bw, y.data_ptr(), rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, bw, y.data_ptr(), rows)
        e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"M": M, "taps_per_band": P, "samples": n, "ms": ms, "MS_per_s": n / ms / 1e3,
                      "frac_of_measured_hbm": 12 * n / (ms * 1e-3) / 6456.2e9}), flush=True)
    ch.close()
