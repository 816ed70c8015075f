"""Multi-GPU host logic on CPU: the time-shard planner (taps-1 halo, frame/rotation aligned) and a
world_size-2 gloo run in which each rank channelizes its shard (with the oracle standing in for the
GPU kernel) and rank 0 stitches the rows in time order — bit-identical to the unsharded run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdr_channelizer_b200.sharding import plan_time_shards, stitch_rows
from tests import synth


@pytest.mark.parametrize("M,P,os_,n,world", [(64, 16, 1, 64 * 1000 + 5, 2), (64, 16, 1, 64 * 1000, 8), (1024, 16, 2, 1024 * 77, 4),
                                             (4096, 16, 1, 4096 * 40 + 100, 8), (8, 8, 1, 1000, 3), (256, 12, 2, 256 * 9, 8)])
def test_plan_covers_rows_once_with_aligned_halo(M, P, os_, n, world):
    D, L = M // os_, M * P
    shards = plan_time_shards(n, M, L, os_, world)
    assert len(shards) == world
    assert shards[0].row_begin == 0 and shards[-1].row_end == n // D
    for a, b in zip(shards, shards[1:]):
        assert a.row_end == b.row_begin                      # contiguous, no gap, no overlap
    for s in shards:
        assert s.sample_begin % (2 * M) == 0                 # rotation restarts in phase, row parity preserved
        assert s.sample_end == s.row_end * D <= n
        if s.rows:
            assert s.sample_begin <= max(0, s.row_begin * D - (L - 1))   # halo of taps-1 samples
            assert s.row_begin * D - s.sample_begin < L - 1 + 2 * M      # ... and not much more
        assert s.discard_rows == s.row_begin - s.sample_begin // D
    sizes = [s.rows for s in shards]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, iq, bw, M, P, os_, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as orc
    taps = orc.design_prototype(M, P)
    shard = plan_time_shards(len(iq), M, M * P, os_, world)[rank]
    part = orc.channelize_raw(iq[shard.sample_begin:shard.sample_end], bw, M, taps, os_)[shard.discard_rows:]
    assert part.shape[0] == shard.rows
    # the only exchange on this path: rank 0 collects row blocks in time order (no data-path collective)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(part, gathered, dst=0)
    # max-over-ranks reduction as bench.py does for its timing
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        np.save(out_path, stitch_rows(gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("M,P,os_", [(64, 16, 1), (32, 12, 2)])
def test_two_rank_gloo_sharded_equals_unsharded(tmp_path, orc, M, P, os_):
    n = M * 400 + 7
    iq, bw = synth.tones_int16_q11(n, M, seed=6)
    out_path = str(tmp_path / "stitched.npy")
    mp.spawn(_worker, args=(2, _free_port(), iq, bw, M, P, os_, out_path), nprocs=2, join=True)
    whole = orc.channelize_raw(iq, bw, M, orc.design_prototype(M, P), os_)
    got = np.load(out_path)
    assert got.shape == whole.shape and np.array_equal(got, whole)


def _gather_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdr_channelizer_b200.sharding import gather_rows_to_rank
    rows = [5, 4, 4][:world]
    g = torch.Generator().manual_seed(rank)
    y = torch.complex(torch.randn(rows[rank], 8, generator=g), torch.randn(rows[rank], 8, generator=g))
    full = gather_rows_to_rank(y, rows, dst=0)
    if rank == 0:
        torch.save(full, out_path)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_in_time_order_gloo(tmp_path):
    """The PDW stage's one exchange: shard row blocks collected on rank 0 in rank (= time) order."""
    out_path = str(tmp_path / "full.pt")
    mp.spawn(_gather_worker, args=(3, _free_port(), out_path), nprocs=3, join=True)
    full = torch.load(out_path)
    expect = []
    for r, n in enumerate([5, 4, 4]):
        g = torch.Generator().manual_seed(r)
        expect.append(torch.complex(torch.randn(n, 8, generator=g), torch.randn(n, 8, generator=g)))
    assert torch.equal(full, torch.cat(expect))


# ---- PDW extraction over time shards: distributed median + boundary stitching -------------
def _pdw_matrix(rows=600, M=8, seed=5):
    """A channel matrix with pulses placed to exercise every boundary case of a 3-way row split
    (200 rows each): inside one shard, straddling one boundary, spanning a whole shard, ending exactly
    on a shard's last row, starting on a shard's first row, and open at the end (dropped, :135)."""
    rng = np.random.default_rng(seed)
    y = (rng.normal(size=(rows, M)) + 1j * rng.normal(size=(rows, M))) * 0.01
    def pulse(k, a, b, f=0.05, amp=1.0):
        n = np.arange(b - a)
        y[a:b, k] += amp * np.exp(2j * np.pi * f * n)
    pulse(0, 20, 60)
    pulse(1, 150, 260, f=-0.11)            # straddles boundary 200
    pulse(2, 180, 450, f=0.2, amp=0.95)    # spans the whole middle shard
    pulse(3, 100, 200)                     # last pulse sample is the last row of shard 0
    pulse(4, 200, 230)                     # first pulse sample is the first row of shard 1
    pulse(5, 390, 410); pulse(5, 30, 50)   # two pulses in one channel, one straddling boundary 400
    pulse(6, 580, 600)                     # still open at the end of the recording
    return y.astype(np.complex64)


def _pdw_worker(rank, world, port, y, bounds, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdr_channelizer_b200.sharding import TorchDistComm, create_pdws_sharded
    from tests.fake_pdw_shard import FakePdwShard
    a, b = bounds[rank], bounds[rank + 1]
    shard = FakePdwShard(y[a:b], a, y.shape[0], D=8, fs=8e6, fc=1e9, t0=3.0)
    recs, nf = create_pdws_sharded(shard, TorchDistComm())
    if rank == world - 1:
        np.save(out_path, np.array([[r.toa_row, r.end_row, r.channel, r.amp, r.freq_hz, r.snr_db, r.toa_s, r.pw_s, r.saturated]
                                    for r in recs] + [list(nf) + [0.0]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bounds", [[0, 200, 400, 600], [0, 600], [0, 199, 600], [0, 1, 2, 600], [0, 0, 600, 600]])
def test_sharded_pdws_gloo_equal_the_oracle_on_the_stitched_matrix(tmp_path, orc, bounds):
    """World-size 1..3 gloo runs of create_pdws_sharded (numpy stand-in for the GPU stages) against the
    oracle's create_pdws_channelized.m restatement on the whole matrix: same pulses, same order, same rows."""
    y = _pdw_matrix()
    world = len(bounds) - 1
    out_path = str(tmp_path / "recs.npy")
    mp.spawn(_pdw_worker, args=(world, _free_port(), y, bounds, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    ref, nf = orc.pdws(y.astype(np.complex128), D=8, fc_hz=1e9, fs_sps=8e6, t0=3.0)
    assert np.allclose(got[-1][:8], nf, rtol=1e-6)
    got = got[:-1]
    assert len(ref) == len(got) == 7
    for g, r in zip(got, ref):
        assert (int(g[0]), int(g[1]), int(g[2]), int(g[8])) == (r.toa_row, r.end_row, r.channel, r.saturated)
        assert abs(g[3] - r.amp) <= 1e-6 * r.amp and abs(g[5] - r.snr_db) <= 1e-5
        assert abs(g[4] - r.freq_hz) <= 1e-3 * 1e6 / 360 and g[6] == r.toa_s and g[7] == r.pw_s


def test_fold_exit_codes():
    from sdr_channelizer_b200.sharding import fold_exit_codes
    codes = [np.array([0, 1, 2, 3, 1, 0], dtype=np.uint8), np.array([2, 2, 2, 2, 3, 3], dtype=np.uint8)]
    assert fold_exit_codes([], 6).tolist() == [0] * 6
    assert fold_exit_codes(codes[:1], 6).tolist() == [0, 1, 0, 1, 1, 0]
    assert fold_exit_codes(codes, 6).tolist() == [0, 1, 0, 1, 0, 1]
