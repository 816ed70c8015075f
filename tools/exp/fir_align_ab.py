"""Does the one-element misalignment of k_fir's warp loads (branch p reads x[mM - p]: a warp covers indices == 1..32
mod 32) cost DRAM traffic?  Timing-only A/B: the same launch with the input pointer moved by 124 bytes, which makes
every warp's 128 bytes line-aligned (results are then those of a shifted recording)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg
M, P = 4096, 16
for n in (280_000_000 // M * M, 3_686_400_000):
    x = torch.randint(-2048, 2048, (n + 64, 2), dtype=torch.int16, device="cuda")
    rows = n // M
    y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
    st = torch.cuda.current_stream(); ch.set_stream(st.cuda_stream)
    for off in (0, 124):
        for _ in range(2):
            ch.reset(); ch.process_ptr(x.data_ptr() + off, n, 12, y.data_ptr(), rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            ch.reset(); ch.process_ptr(x.data_ptr() + off, n, 12, y.data_ptr(), rows)
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"samples": n, "input_offset_bytes": off, "ms": round(ms, 3), "GS_per_s": round(n / ms / 1e6, 1),
                          "frac": round(12 * n / ms / 1e6 / 6456.2, 4)}), flush=True)
    ch.close(); del x, y; torch.cuda.empty_cache()
