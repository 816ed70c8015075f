"""Time sharding of one recording across GPUs (BASELINE.json configs[3]).

Output row m depends only on samples x[m*D - (L-1) .. m*D], so the recording splits into contiguous
row ranges; shard g reads its own samples plus a halo of L-1 = taps-1 preceding samples (zeros for
shard 0).  Shards are independent: no collective on the channelizer path, rows are stitched in
time order on the host.  Feeding the halo through the same channelizer first (and discarding the
rows it produces) makes the shard's FIR state identical to the single-GPU run, so results are
bit-identical by construction.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class TimeShard:
    rank: int
    row_begin: int        # first output row owned (global index)
    row_end: int          # one past the last owned row
    sample_begin: int     # first sample to feed (includes the halo, frame aligned)
    sample_end: int       # one past the last sample to feed
    discard_rows: int     # rows produced by the halo that belong to the previous shard

    @property
    def rows(self):
        return self.row_end - self.row_begin

    @property
    def samples(self):
        return self.sample_end - self.sample_begin


def plan_time_shards(num_samples, M, ntaps, oversample, world_size):
    """Split floor(num_samples / D) rows into world_size contiguous ranges.

    Each shard starts feeding at a frame boundary at least (taps-1) samples before its first owned
    row's newest sample, rounded down to a multiple of 2*M so the circular branch rotation (m*D mod M)
    restarts in phase and global row parity is preserved.  Rows the halo itself produces are discarded (their FIR history is incomplete).
    """
    D = M // oversample
    total_rows = num_samples // D
    shards = []
    base, extra = divmod(total_rows, world_size)
    row = 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        rb, re = row, row + n
        row = re
        # newest sample of row rb is rb*D; oldest it touches is rb*D - (ntaps-1)
        start = max(0, rb * D - (ntaps - 1))
        # aligned to 2*M samples: frames and the circular rotation (m*D mod M) restart in phase, and a
        # row keeps the parity of its global index, which fixes the tap order the kernel sums it in
        # (rows are filtered in pairs) -> shard results are bit-identical to the unsharded run
        start = (start // (2 * M)) * (2 * M)
        discard = rb - start // D                     # rows start//D .. rb-1 come out of the halo
        shards.append(TimeShard(r, rb, re, start, re * D, discard))
    return shards


def stitch_rows(parts):
    """Concatenate per-shard row blocks (already trimmed of their discard rows) in rank order."""
    import numpy as np
    return np.concatenate(list(parts), axis=0) if parts else None


def gather_rows_to_rank(y_local, rows_per_rank, dst=0, group=None):
    """Multi-GPU PDW stage, simplest exact form: the PDW threshold is the per-channel
    median of |y| over the WHOLE recording (create_pdws_channelized.m:73), so the time shards' channel
    matrices are collected on one GPU over NCCL/NVLink (gather in time order) and chz_pdws_dev runs there
    on the stitched matrix.  y_local: torch complex64 [rows, M] on this rank's GPU; rows_per_rank: list
    of every rank's row count.  Returns the stitched [sum(rows), M] tensor on `dst`, None elsewhere.
    (A distributed median by histogram all-reduce would avoid moving y; not built yet.)"""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    M = y_local.shape[1]
    if rank == dst:
        out = torch.empty((sum(rows_per_rank), M), dtype=y_local.dtype, device=y_local.device)
        parts, pos = [], 0
        for r in rows_per_rank:
            parts.append(out[pos:pos + r]); pos += r
    else:
        out, parts = None, None
    # rows differ by at most one between ranks: exchange as float32 views with point-to-point copies in rank order
    if rank == dst:
        parts[dst].copy_(y_local)
        reqs = [dist.irecv(torch.view_as_real(parts[r]), src=r, group=group) for r in range(world) if r != dst]
        for q in reqs:
            q.wait()
    else:
        dist.send(torch.view_as_real(y_local.contiguous()), dst=dst, group=group)
    return out
