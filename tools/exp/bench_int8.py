"""configs[0] geometry (and M = 64) on 8-bit recordings at a size that can be timed: 560 M int8 samples."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg
for M, P, os_ in ((8, 8, 1), (8, 8, 2), (64, 16, 1), (64, 12, 1), (256, 16, 1)):
    n = 560_000_000 // M * M
    x = torch.randint(-128, 128, (n, 2), dtype=torch.int8, device="cuda")
    rows = n // (M // os_)
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P), OversamplingRatio=os_)
    st = torch.cuda.Stream(); ch.set_stream(st.cuda_stream); torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, 8, y.data_ptr(), rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, 8, y.data_ptr(), rows)
        e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"M": M, "P": P, "oversample": os_, "bit_width": 8, "ms": round(ms, 4), "GS_per_s": round(n / ms / 1e6, 1),
                      "frac_of_measured_hbm": round((2 + 8 * os_) * n / (ms * 1e-3) / 6456.2e9, 4)}), flush=True)
    ch.close(); del x, y; torch.cuda.empty_cache()
