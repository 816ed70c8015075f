#!/usr/bin/env python
"""BASELINE.json configs[3]: 4096-channel channelizer, 65536-tap prototype, ONE 60 s recording at 61.44 MS/s
(3 686.4 M samples) time-sharded across the ranks (strong scaling), each shard carrying its taps-1 halo.
torchrun --nproc-per-node N tools/bench_cfg4_dist.py ; rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sdr_channelizer_b200 as pkg  # noqa: E402

M, P, OS, BW = 4096, 16, 1, 12
TOTAL = 61_440_000 * 60
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
sh = pkg.plan_time_shards(TOTAL, M, M * P, OS, world)[rank]
g = torch.Generator(device=dev).manual_seed(4 + rank)
x = torch.randint(-2048, 2048, (sh.samples, 2), dtype=torch.int16, device=dev, generator=g)
rows = sh.samples // M
y = torch.empty((rows, M), dtype=torch.complex64, device=dev)
ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ch.set_stream(st.cuda_stream)
steps = 5
for _ in range(3):
    ch.reset(); ch.process_ptr(x.data_ptr(), sh.samples, BW, y.data_ptr(), rows)
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    ch.reset(); ch.process_ptr(x.data_ptr(), sh.samples, BW, y.data_ptr(), rows)
e1.record(st)
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if dist is not None:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    peak = 6456.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({"config": "configs[3]", "n_gpus": world, "total_samples": TOTAL, "halo_samples": M * P - 1,
                      "ms_per_pass_max_over_ranks": ms, "aggregate_MS_per_s": TOTAL / ms / 1e3,
                      "frac_of_hbm_roofline_per_gpu": TOTAL / world * 12 / (ms * 1e-3) / 1e9 / peak, "scaling": "strong"}), flush=True)
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
ch.close()
