#!/usr/bin/env python
"""BASELINE.json configs[4] across GPUs: generate_channelized_training_iq-style pulsed files (100 ms @
56 MS/s, int16) -> 256 channels -> PDWs.  Files are independent (the reference computes the noise floor
per file, create_pdws_channelized.m:22-27,73), so ranks take files round-robin: replicas, no exchange
on the data path.  Launch with torchrun (or plain python for one GPU); prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sdr_channelizer_b200 as pkg  # noqa: E402

M, P, FS = 256, 16, 56e6
N = 5_600_000 // M * M
FILES_PER_RANK = int(os.environ.get("FILES_PER_RANK", "16"))


def make_file(seed, dev):
    """Device-side version of tests/synth.pulsed_int16 (same recipe; content only steers the PDW count)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    u = torch.rand(4, generator=g, device=dev).tolist()
    f = -(FS / 2) + FS * u[0]
    pw = 10e-6 + 990e-6 * u[1]
    pri = max(10e-6, pw) + (10000e-6 - max(10e-6, pw)) * u[2]
    npw, npri = int(round(FS * pw)), int(round(FS * pri))
    start = int(u[3] * npri)
    t = torch.arange(N, device=dev, dtype=torch.float64)
    rel = torch.remainder(t - start, npri)
    on = (t >= start) & (rel < npw)
    ph = 2 * torch.pi * torch.frac(rel * (f / FS))
    re = torch.randn(N, device=dev, generator=g) * 0.005 + 0.5 * torch.cos(ph).float() * on
    im = torch.randn(N, device=dev, generator=g) * 0.005 + 0.5 * torch.sin(ph).float() * on
    return torch.stack([torch.clamp(torch.round(re * 32768), -32768, 32767), torch.clamp(torch.round(im * 32768), -32768, 32767)],
                       dim=1).to(torch.int16).contiguous()


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    files = [make_file(1000 + rank + world * i, dev) for i in range(FILES_PER_RANK)]
    rows = N // M
    workers = int(os.environ.get("WORKERS", "4"))       # host threads per GPU, one handle + stream each: the
    import threading                                      # per-file kernels are small, several files overlap
    taps = pkg.design_prototype(M, P)
    ctx = []
    for w in range(workers):
        ch = pkg.Channelizer(M, taps=taps)
        st = torch.cuda.Stream(device=dev)
        ch.set_stream(st.cuda_stream)
        ctx.append((ch, torch.empty((rows, M), dtype=torch.complex64, device=dev)))
    torch.cuda.synchronize()

    def one(w, x):
        ch, y = ctx[w]
        ch.reset()
        ch.process_ptr(x.data_ptr(), N, 16, y.data_ptr(), rows)
        recs, _ = ch.pdws_ptr(y.data_ptr(), rows, FS, 2.4e9, 0.0)
        return len(recs)

    counts = [0] * workers

    def work(w):
        torch.cuda.set_device(local)
        for i in range(w, len(files), workers):
            counts[w] += one(w, files[i])

    for w in range(workers):
        one(w, files[0])
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    threads = [threading.Thread(target=work, args=(w,)) for w in range(workers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    npdw = sum(counts)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    cnt = torch.tensor([npdw], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        total = FILES_PER_RANK * world
        print(json.dumps({"config": "configs[4]", "n_gpus": world, "files": total, "samples_per_file": N, "pdws": int(cnt.item()),
                          "seconds": float(dt.item()), "files_per_s": total / float(dt.item()),
                          "input_MS_per_s": total * N / float(dt.item()) / 1e6, "pdws_per_s": cnt.item() / float(dt.item()),
                          "parallelism": f"file replicas round-robin, no collective; {workers} host threads/handles per GPU"}), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    for ch, _ in ctx:
        ch.close()


if __name__ == "__main__":
    main()
