"""K3 alone (row FFT of a device-resident matrix, chz_fft_rows_dev) for M = 64: the CUDA-core kernel the tensor-core
DFT-as-GEMM variant (tools/ubench/dft64_tc.cu) is compared with."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg
M = 64
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 9_600_000
u = torch.randn((rows, M), dtype=torch.complex64, device="cuda")
y = torch.empty_like(u)
ch = pkg.Channelizer(M, NumTapsPerBand=8)
st = torch.cuda.current_stream(); ch.set_stream(st.cuda_stream)
for _ in range(2):
    ch.fft_rows_ptr(u.data_ptr(), y.data_ptr(), rows)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(5):
    ch.fft_rows_ptr(u.data_ptr(), y.data_ptr(), rows)
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
ref = torch.fft.ifft(u[:300].to(torch.complex128), dim=1) * M
err = (torch.linalg.norm(y[:300].to(torch.complex128) - ref) / torch.linalg.norm(ref)).item()
print(json.dumps({"kernel": "k_fft_rows<64> (CUDA cores, radix 8x8 in shared memory)", "rows": rows, "ms": round(ms, 4),
                  "GS_per_s": round(rows * M / ms / 1e6, 1), "GBps_16B_per_sample": round(rows * M * 16 / ms / 1e6, 1), "rel_rms": err}))
