"""Does the split path's FIR output stay in L2 when the recording is processed in row chunks?
usage: chunked_split.py M P chunk_rows [samples]   (chunk_rows = 0: one call).  Streams the chunks through one
stateful handle (CHZ_OPT_FORCE_PATH = 2) and prints time and GS/s; run under `ncu --metrics dram__bytes_*` for bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg

M, P, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = (int(sys.argv[4]) if len(sys.argv) > 4 else 280_000_000) // M * M
steps = int(os.environ.get("STEPS", "3"))
x = torch.randint(-2048, 2048, (n, 2), dtype=torch.int16, device="cuda")
rows = n // M
y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
ch.set_option(pkg.CHZ_OPT_FORCE_PATH, int(os.environ.get("PATH_ID", "2")))
st = torch.cuda.Stream(); ch.set_stream(st.cuda_stream); torch.cuda.synchronize()
C = C or rows

def run():
    ch.reset()
    r = 0
    while r < rows:
        c = min(C, rows - r)
        ch.process_ptr(x.data_ptr() + r * M * 4, c * M, 16, y.data_ptr() + r * M * 8, c)
        r += c

with torch.cuda.stream(st):
    run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        run()
    e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"M": M, "P": P, "chunk_rows": C, "samples": n, "ms": round(ms, 4), "GS_per_s": round(n / ms / 1e6, 1),
                  "frac_of_hbm_12B": round(12 * n / (ms * 1e-3) / 6456.2e9, 4), "launches": ch.launches if hasattr(ch, "launches") else None}), flush=True)
