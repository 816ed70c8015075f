"""A/B of the large-M paths (CHZ_OPT_FORCE_PATH) on one B200: split (2), L2-ring cluster kernels (3, 7, 8, 9),
pipelined task queue (6), DSMEM st.async (10).  One JSON line per (M, oversample, path)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg

paths = [int(p) for p in (sys.argv[1].split(",") if len(sys.argv) > 1 else "2,3,6,7,8,9,10".split(","))]
for M, os_ in ((1024, 2), (1024, 1), (2048, 1), (2048, 2), (4096, 1)):
    n = 280_000_000 // M * M
    x = torch.randint(-2000, 2000, (n, 2), dtype=torch.int16, device="cuda")
    rows = n // (M // os_)
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    for path in paths:
        ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, 16), OversamplingRatio=os_)
        ch.set_option(pkg.CHZ_OPT_FORCE_PATH, path)
        st = torch.cuda.Stream(); ch.set_stream(st.cuda_stream); torch.cuda.synchronize()
        with torch.cuda.stream(st):
            for _ in range(2):
                ch.reset(); ch.process_ptr(x.data_ptr(), n, 16, y.data_ptr(), rows)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5):
                ch.reset(); ch.process_ptr(x.data_ptr(), n, 16, y.data_ptr(), rows)
            e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"M": M, "oversample": os_, "path": path, "ms": round(ms, 4), "GS_per_s": round(n / ms / 1e6, 1),
                          "frac_of_measured_hbm": round((4 + 8 * os_) * n / (ms * 1e-3) / 6456.2e9, 4)}), flush=True)
        ch.close()
    del x, y
    torch.cuda.empty_cache()
