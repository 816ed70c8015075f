#!/usr/bin/env python
"""Device-resident throughput of every BASELINE.json config (bench.py measures only the headline
configs[1]).  One JSON line per config: input MS/s, ms per pass, algorithmic GB/s and the fraction of
the measured HBM peak.  Run on a B200:  python tools/bench_configs.py [--only cfg2,cfg4] [--scale 0.25]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import sdr_channelizer_b200 as pkg  # noqa: E402
from tests import synth  # noqa: E402

CONFIGS = {
    # name: (M, taps/band, oversample, bit width, samples, note)
    "cfg1": (8, 8, 1, 8, 1_000_000, "configs[0]: 8 ch, 64 taps, 1M-sample 8-bit file"),
    "cfg2": (64, 16, 1, 12, 614_400_000, "configs[1]: 64 ch, 1024 taps, 61.44 MS/s x 10 s, 12-bit"),
    "cfg3": (1024, 16, 2, 16, 560_000_000, "configs[2]: 1024 ch 2x oversampled, 16384 taps, 56 MS/s x 10 s"),
    "cfg4": (4096, 16, 1, 12, 3_686_400_000, "configs[3]: 4096 ch, 65536 taps, 61.44 MS/s x 60 s (one GPU's view)"),
    "cfg5chan": (256, 16, 1, 16, 5_600_000 * 8, "configs[4] channelizer part: 256 ch, 8 files x 100 ms @ 56 MS/s back to back"),
}


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run_chan(name, scale, steps):
    M, P, os_, bw, n, note = CONFIGS[name]
    n = int(n * scale) // M * M
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    lim = 2 ** (bw - 1)
    if bw <= 8:
        x = torch.randint(-lim, lim, (n, 2), dtype=torch.int8, device=dev, generator=g)
    else:
        x = torch.randint(-lim, lim, (n, 2), dtype=torch.int16, device=dev, generator=g)
    rows = n // (M // os_)
    y = torch.empty((rows, M), dtype=torch.complex64, device=dev)
    ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P), OversamplingRatio=os_)
    if os.environ.get("CHZ_BENCH_PATH"):          # kernel A/B experiments
        ch.set_option(pkg.CHZ_OPT_FORCE_PATH, int(os.environ["CHZ_BENCH_PATH"]))
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        ch.set_stream(st.cuda_stream)
        for _ in range(3):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, bw, y.data_ptr(), rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ch.kernel_launches
        e0.record(st)
        for _ in range(steps):
            ch.reset(); ch.process_ptr(x.data_ptr(), n, bw, y.data_ptr(), rows)
        e1.record(st)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    bytes_algo = n * ((2 if bw <= 8 else 4) + 8 * os_)
    gbs = bytes_algo / (ms * 1e-3) / 1e9
    out = {"config": name, "note": note, "M": M, "taps": M * P, "oversample": os_, "bit_width": bw, "samples": n,
           "ms_per_pass": ms, "MS_per_s": n / (ms * 1e-3) / 1e6, "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peak(),
           "launches_per_pass": (ch.kernel_launches - l0) / steps}
    ch.close()
    del x, y
    torch.cuda.empty_cache()
    return out


def run_pdw(files=8):
    """configs[4]: pulsed files -> 256 channels -> PDWs; reports the device time of each stage."""
    M, P = 256, 16
    fs = 56e6
    n = 5_600_000 // M * M
    taps = pkg.design_prototype(M, P)
    tot_chan = tot_pdw = 0.0
    npdw = 0
    rows = n // M
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(M, taps=taps)          # one handle, reset per file (a fresh object per file in the reference)
    st = torch.cuda.Stream()
    ch.set_stream(st.cuda_stream)
    for i in range(-1, files):                  # file -1 is an untimed warm-up
        iq, bw, _ = synth.pulsed_int16(n, M=M, seed=100 + max(i, 0), fs=fs)
        d_in = torch.from_numpy(iq).cuda()
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            ch.reset()
            t0 = time.perf_counter()
            ch.process_ptr(d_in.data_ptr(), n, bw, y.data_ptr(), rows); torch.cuda.synchronize()
            t1 = time.perf_counter()
            recs, _ = ch.pdws_ptr(y.data_ptr(), rows, fs)
            t2 = time.perf_counter()
        if i >= 0:
            tot_chan += t1 - t0; tot_pdw += t2 - t1; npdw += len(recs)
    ch.close()
    return {"config": "cfg5", "note": "configs[4]: 8 pulsed files (100 ms @ 56 MS/s, int16) -> 256 ch -> PDWs, one GPU, host-timed per file",
            "files": files, "samples_per_file": n, "pdws": npdw, "chan_ms_per_file": tot_chan / files * 1e3,
            "pdw_ms_per_file": tot_pdw / files * 1e3, "MS_per_s_end_to_end_device": files * n / (tot_chan + tot_pdw) / 1e6,
            "pdws_per_s": npdw / (tot_chan + tot_pdw)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    names = [s for s in a.only.split(",") if s] or list(CONFIGS) + ["cfg5"]
    for nm in names:
        try:
            r = run_pdw() if nm == "cfg5" else run_chan(nm, a.scale if nm != "cfg1" else 1.0, a.steps)
        except Exception as e:   # keep going: one config must not hide the others
            r = {"config": nm, "error": repr(e)}
        print(json.dumps(r), flush=True)
