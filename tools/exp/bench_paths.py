"""A/B of channelizer kernel paths on one B200: one JSON line per (M, oversample, taps/band, bits, path).
usage: bench_paths.py "M,os,P,bits,path[,samples]" ...      (path = CHZ_OPT_FORCE_PATH value, 0 = default)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg

PEAK = 6456.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
steps = int(os.environ.get("STEPS", "5"))
for spec in sys.argv[1:]:
    f = [int(v) for v in spec.split(",")]
    M, os_, P, bits, path = f[:5]
    n = (f[5] if len(f) > 5 else 280_000_000) // M * M
    lim = 2 ** (bits - 1)
    x = torch.randint(-lim, lim, (n, 2), dtype=torch.int8 if bits <= 8 else torch.int16, device="cuda")
    rows = n // (M // os_)
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P), OversamplingRatio=os_)
    try:
        ch.set_option(pkg.CHZ_OPT_FORCE_PATH, path)
    except Exception as e:
        print(json.dumps({"spec": spec, "error": repr(e)}), flush=True)
        continue
    st = torch.cuda.Stream(); ch.set_stream(st.cuda_stream); torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, bits, y.data_ptr(), rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, bits, y.data_ptr(), rows)
        e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    bps = (2 if bits <= 8 else 4) + 8 * os_
    print(json.dumps({"M": M, "oversample": os_, "P": P, "bits": bits, "path": path, "samples": n, "ms": round(ms, 4),
                      "GS_per_s": round(n / ms / 1e6, 1), "frac_of_measured_hbm": round(bps * n / (ms * 1e-3) / (PEAK * 1e9), 4),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("CHZ_")}}), flush=True)
    ch.close()
    del x, y
    torch.cuda.empty_cache()
