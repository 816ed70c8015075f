#!/usr/bin/env python
"""north_star: "Tensor cores are evaluated only for a DFT-as-GEMM variant at small M, and kept only if
ncu shows it winning."  This measures the best case for that variant on the headline geometry
(M = 64): the M-point DFT of every row as a dense GEMM Y[rows, 2M] = U[rows, 2M] x W[2M, 2M] (complex
as 2x2 real blocks) on the tensor cores through cuBLAS, in TF32 (10-bit mantissa: NOT accurate enough
for the 1e-5 budget), 3xTF32-equivalent cost (x3) and BF16x3-style split cost, against the time of
our whole fused unpack+FIR+FFT kernel on the same number of rows.  The GEMM alone reads and writes
the fp32 branch matrix once (16 B per sample), i.e. it can only replace the FFT *stage*, and would
still need the FIR output to pass through shared memory / TMEM in a fused kernel.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdr_channelizer_b200 as pkg  # noqa: E402

M, P = 64, 16
rows = 614_400_000 // M // 4          # a quarter of configs[1] (1.2 GB in + 1.2 GB out as real [rows, 128])
dev = torch.device("cuda")
u = torch.randn(rows, 2 * M, device=dev)
k = torch.arange(M, device=dev, dtype=torch.float64)
ang = 2 * torch.pi * torch.outer(k, k) / M
Wc, Ws = torch.cos(ang), torch.sin(ang)
W = torch.zeros(2 * M, 2 * M, dtype=torch.float64, device=dev)
W[0::2, 0::2], W[1::2, 0::2], W[0::2, 1::2], W[1::2, 1::2] = Wc, -Ws, Ws, Wc   # (re,im) interleaved complex product
W = W.float()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = torch.empty_like(u)
res = {"rows": rows, "M": M}
torch.backends.cuda.matmul.allow_tf32 = True
res["gemm_tf32_ms"] = timeit(lambda: torch.matmul(u, W, out=out))
ref = torch.matmul(u[:4096].double(), W.double())
res["gemm_tf32_rel_rms"] = float(((out[:4096].double() - ref).norm() / ref.norm()).item())
torch.backends.cuda.matmul.allow_tf32 = False
res["gemm_fp32_ms"] = timeit(lambda: torch.matmul(u, W, out=out))
res["gemm_fp32_rel_rms"] = float(((out[:4096].double() - ref).norm() / ref.norm()).item())
ub, Wb = u.bfloat16(), W.bfloat16()
outb = torch.empty(rows, 2 * M, dtype=torch.bfloat16, device=dev)
res["gemm_bf16_ms"] = timeit(lambda: torch.matmul(ub, Wb, out=outb))

# our whole fused kernel on the same number of output rows
n = rows * M
x = torch.randint(-2048, 2048, (n, 2), dtype=torch.int16, device=dev)
y = torch.empty((rows, M), dtype=torch.complex64, device=dev)
ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
st = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(st):
    ch.set_stream(st.cuda_stream)

    def fused():
        ch.reset()
        ch.process_ptr(x.data_ptr(), n, 12, y.data_ptr(), rows)
    for _ in range(3):
        fused()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        fused()
    e1.record(st)
    torch.cuda.synchronize()
res["fused_unpack_fir_fft_ms"] = e0.elapsed_time(e1) / 10
res["verdict"] = ("DFT-as-GEMM (FFT stage only, TF32, inaccurate) takes %.2fx the time of the whole fused kernel; "
                  "3xTF32 for 1e-5 accuracy would take ~%.2fx" % (res["gemm_tf32_ms"] / res["fused_unpack_fir_fft_ms"],
                                                                   3 * res["gemm_tf32_ms"] / res["fused_unpack_fir_fft_ms"]))
print(json.dumps(res))
