#!/usr/bin/env python
"""Time-sharded channelizer + PDW extraction across GPUs, checked against the one-GPU run.
torchrun --nproc-per-node N tools/run_sharded_pdw.py   (rank 0 prints one JSON line)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import sdr_channelizer_b200 as pkg  # noqa: E402
from sdr_channelizer_b200.sharding import PdwShard, TorchDistComm, create_pdws_sharded, gather_rows_to_rank  # noqa: E402
from tests import synth  # noqa: E402

M, P, OS = 64, 16, 1
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = M * 200_000
iq, bw, fs = synth.pulsed_int16(n, M=M, seed=4242)          # every rank builds the same recording
# bursts centred on the quarter points, so pulses straddle the shard boundaries of 2- and 4-GPU runs
for q, f in ((0.25, 0.11), (0.5, -0.23), (0.75, 0.37)):
    a, b = int(q * n) - 15000, int(q * n) + 22000
    t = np.arange(b - a)
    burst = 9000.0 * np.exp(2j * np.pi * f * t)
    iq[a:b, 0] = np.clip(iq[a:b, 0] + np.round(burst.real), -32768, 32767).astype(iq.dtype)
    iq[a:b, 1] = np.clip(iq[a:b, 1] + np.round(burst.imag), -32768, 32767).astype(iq.dtype)
taps = pkg.design_prototype(M, P)
shards = pkg.plan_time_shards(n, M, M * P, OS, world)
sh = shards[rank]
ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=OS)
x = torch.from_numpy(iq[sh.sample_begin:sh.sample_end]).to(dev)
rows_all = x.shape[0] // (M // OS)
y_all = torch.empty((rows_all, M), dtype=torch.complex64, device=dev)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
ch.process_ptr(x.data_ptr(), x.shape[0], bw, y_all.data_ptr(), rows_all); ch.synchronize()
y_own = y_all[sh.discard_rows:]                              # halo rows dropped
full = gather_rows_to_rank(y_own, [s.rows for s in shards], dst=0)
recs = None
if rank == 0:
    recs, nf = ch.pdws_ptr(full.data_ptr(), full.shape[0], fs, 2.4e9, 0.0)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
# the same without moving y: distributed median (histogram all-reduce over NCCL) + boundary stitching
comm = TorchDistComm()
shard = PdwShard(ch, y_own.data_ptr(), y_own.shape[0], sh.row_begin, n // (M // OS), fs, 2.4e9, 0.0)
drecs, dnf = create_pdws_sharded(shard, comm)                 # warm-up (NCCL communicator set-up)
torch.cuda.synchronize(); dist.barrier()
t1 = time.perf_counter()
drecs, dnf = create_pdws_sharded(shard, comm)
torch.cuda.synchronize(); dist.barrier()
dt_dist = time.perf_counter() - t1
if rank == 0:
    ch.reset()
    xs = torch.from_numpy(iq).to(dev)
    ys = torch.empty((n // M, M), dtype=torch.complex64, device=dev)
    ch.process_ptr(xs.data_ptr(), n, bw, ys.data_ptr(), n // M); ch.synchronize()
    ref, nf1 = ch.pdws_ptr(ys.data_ptr(), n // M, fs, 2.4e9, 0.0)
    same_y = bool(torch.equal(full.view(torch.float32), ys.view(torch.float32)))
    same_pdw = len(ref) == len(recs) and all((a.channel, a.toa_row, a.end_row, a.amp, a.freq_hz, a.saturated) ==
                                             (b.channel, b.toa_row, b.end_row, b.amp, b.freq_hz, b.saturated) for a, b in zip(recs, ref))
    same_dist = len(ref) == len(drecs) and all(bytes(a) == bytes(b) for a, b in zip(drecs, ref)) and bool(np.array_equal(dnf, nf1))
    straddlers = sum(1 for r in ref if any(s_.row_begin < r.end_row and r.toa_row <= s_.row_begin for s_ in shards[1:]))
    print(json.dumps({"n_gpus": world, "samples": n, "rows": int(full.shape[0]), "pdws": len(recs), "seconds": dt,
                      "stitched_rows_bit_identical_to_one_gpu": same_y, "pdws_identical_to_one_gpu": bool(same_pdw),
                      "distributed_pdws_byte_identical_to_one_gpu": bool(same_dist), "distributed_pdw_seconds": dt_dist,
                      "pulses_straddling_a_shard_boundary": straddlers}), flush=True)
dist.barrier(); dist.destroy_process_group()
ch.close()
