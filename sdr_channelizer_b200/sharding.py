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
    (create_pdws_sharded below does the same without moving y: distributed median + edge stitching.)"""
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


# ------------------------------------------------------------------------------------------------
# PDW extraction over time shards without moving y: distributed exact median by summing
# the radix-select histograms of all shards, FSM entry states folded across shard boundaries, pulses
# that straddle a boundary stitched from a few column segments.  Identical to chz_pdws_dev on the
# stitched matrix.
# ------------------------------------------------------------------------------------------------
class _DevArray:
    """Zero-copy window on device memory for torch.as_tensor (the CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def fold_exit_codes(codes_before, M):
    """State of the edge FSM (create_pdws_channelized.m:83-96) per natural channel on entry to a shard,
    from the exit codes (chz_pdw_shard_exit_state_dev) of the shards before it, in time order:
    0 inactive, 1 active, 2 keeps, 3 toggles the state the shard was entered with."""
    import numpy as np
    state = np.zeros(M, dtype=np.uint8)                       # the FSM starts inactive (:83)
    for c in codes_before:
        c = np.asarray(c, dtype=np.uint8)
        state = np.where(c == 0, 0, np.where(c == 1, 1, np.where(c == 2, state, 1 - state))).astype(np.uint8)
    return state


def pair_events(events, M, reproduce_phase_bug=False):
    """All shards' edge events -> pulses in the reference's order (host only; chz_pdw_pair_events)."""
    import ctypes as C
    import numpy as np
    from . import _lib
    ev = np.ascontiguousarray(events, dtype=np.uint64).copy()
    n = C.c_uint64(0)
    cap = max(len(ev) // 2, 1)
    arr = (_lib.Pulse * cap)()
    _lib.check(_lib.lib().chz_pdw_pair_events(ev.ctypes.data_as(C.c_void_p), len(ev), int(M), int(bool(reproduce_phase_bug)),
                                             C.cast(arr, C.c_void_p), cap, C.byref(n)), "chz_pdw_pair_events")
    return [arr[i] for i in range(int(n.value))]


class PdwShard:
    """The staged extractor (include/channelizer.h, chz_pdw_shard_*) over this rank's rows
    y[nrows][M] (device pointer), which are rows row_offset+1 .. row_offset+nrows of the recording."""

    def __init__(self, channelizer, y_ptr, nrows, row_offset, total_rows, fs, fc=0.0, sampleStartTime=0.0,
                 SNR_THRESHOLD=15.0, sat_level=0.9999, reproduce_phase_bug=False, TRAILING_EDGE_THRESHOLD=None):
        self.ch, self.y_ptr, self.nrows = channelizer, int(y_ptr), int(nrows)
        self.row_offset, self.total_rows = int(row_offset), int(total_rows)
        self.M = channelizer.NumFrequencyBands
        self.bug = bool(reproduce_phase_bug)
        self.params = channelizer._params(fs, fc, sampleStartTime, SNR_THRESHOLD, sat_level, reproduce_phase_bug,
                                          TRAILING_EDGE_THRESHOLD)

    def hist(self, p):
        import ctypes as C
        import torch
        from . import _lib
        ptr, words = C.c_void_p(0), C.c_uint64(0)
        _lib.check(_lib.lib().chz_pdw_shard_hist_dev(self.ch.handle, C.c_void_p(self.y_ptr), self.nrows, p, C.byref(ptr),
                                                    C.byref(words)), "chz_pdw_shard_hist_dev")
        return torch.as_tensor(_DevArray(ptr.value, (int(words.value),), "<i4"), device="cuda")

    def select(self, p):
        from . import _lib
        _lib.check(_lib.lib().chz_pdw_shard_select(self.ch.handle, p, self.total_rows), "chz_pdw_shard_select")

    def thresholds(self):
        import ctypes as C
        from . import _lib
        _lib.check(_lib.lib().chz_pdw_shard_thresholds(self.ch.handle, C.byref(self.params)), "chz_pdw_shard_thresholds")

    def set_noise_floor(self, nf):
        """Skip the median passes: thresholds from a noise floor the caller supplies (chz_pdw_shard_set_noise_floor)."""
        import ctypes as C
        import numpy as np
        from . import _lib
        nf = np.ascontiguousarray(nf, dtype=np.float64)
        assert nf.shape == (self.M,)
        _lib.check(_lib.lib().chz_pdw_shard_set_noise_floor(self.ch.handle, C.byref(self.params), nf.ctypes.data_as(C.c_void_p)),
                   "chz_pdw_shard_set_noise_floor")

    def noise_floor(self):
        import ctypes as C
        import numpy as np
        from . import _lib
        nf = np.empty(self.M, dtype=np.float64)
        _lib.check(_lib.lib().chz_pdw_noise_floor(self.ch.handle, nf.ctypes.data_as(C.c_void_p), self.M), "chz_pdw_noise_floor")
        return nf

    def exit_state(self):
        import ctypes as C
        import numpy as np
        from . import _lib
        code = np.zeros(self.M, dtype=np.uint8)
        _lib.check(_lib.lib().chz_pdw_shard_exit_state_dev(self.ch.handle, C.c_void_p(self.y_ptr), self.nrows,
                                                          code.ctypes.data_as(C.c_void_p)), "chz_pdw_shard_exit_state_dev")
        return code

    def detect(self, entry):
        import ctypes as C
        import numpy as np
        from . import _lib
        entry = np.ascontiguousarray(entry, dtype=np.uint8)
        n = C.c_uint64(0)
        cap = 1 << 16
        while True:
            ev = np.empty(cap, dtype=np.uint64)
            rc = _lib.lib().chz_pdw_shard_detect_dev(self.ch.handle, C.c_void_p(self.y_ptr), self.nrows, self.row_offset,
                                                    entry.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p), cap, C.byref(n))
            if rc == _lib.CHZ_ECAPACITY:
                cap = int(n.value)
                continue
            _lib.check(rc, "chz_pdw_shard_detect_dev")
            return ev[:int(n.value)].copy()

    def _records(self, mat_ptr, ld, row_offset, pulses):
        import ctypes as C
        from . import _lib
        if not pulses:
            return []
        arr = (_lib.Pulse * len(pulses))(*pulses)
        out = (_lib.Pdw * len(pulses))()
        _lib.check(_lib.lib().chz_pdw_shard_records_dev(self.ch.handle, C.byref(self.params), C.c_void_p(mat_ptr), int(ld),
                                                       int(row_offset), C.cast(arr, C.c_void_p), len(pulses),
                                                       C.cast(out, C.c_void_p)), "chz_pdw_shard_records_dev")
        return [bytes(out[i]) for i in range(len(pulses))]

    def records(self, pulses):
        """Records (as bytes of chz_pdw_t) of pulses that lie entirely inside this shard."""
        return self._records(self.y_ptr, self.M, self.row_offset, pulses)

    def column_segment(self, cols, row_lo, row_hi):
        """y[row_lo..row_hi (1-based rows of the recording, inclusive), cols] -> host complex64 [n, len(cols)]."""
        import torch
        y = torch.as_tensor(_DevArray(self.y_ptr, (self.nrows, self.M, 2), "<f4"), device="cuda")
        a, b = row_lo - 1 - self.row_offset, row_hi - self.row_offset
        seg = y[a:b][:, list(cols)].contiguous().cpu().numpy()
        return seg.view("<c8")[..., 0]

    def records_from_matrix(self, mat, row_offset, pulses):
        """Records of pulses whose rows are those of the small host matrix `mat` (complex64 [n, w], first row =
        row row_offset+1 of the recording); pulses carry columns of `mat` in col / col_phase."""
        import numpy as np
        import torch
        d = torch.from_numpy(np.ascontiguousarray(mat, dtype=np.complex64)).cuda()
        torch.cuda.synchronize()
        return self._records(d.data_ptr(), mat.shape[1], row_offset, pulses)


class TorchDistComm:
    """The exchanges of create_pdws_sharded over torch.distributed (NCCL between GPUs; gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce_sum_(self, t):
        import torch
        import torch.distributed as dist
        dist.all_reduce(t, group=self.group)
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()   # the library works on its own stream

    def all_gather(self, obj):
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def all_gather_array(self, arr):
        """Variable-length 1-D numpy arrays of one dtype -> list of the ranks' arrays, as two tensor collectives (the
        lengths, then the zero-padded payloads as bytes) instead of pickled objects: with NCCL the bytes travel GPU to
        GPU over NVLink and nothing is serialised (all_gather_object was the dominant cost of the 8-GPU extraction)."""
        import numpy as np
        import torch
        import torch.distributed as dist
        arr = np.ascontiguousarray(arr)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        nbytes = torch.tensor([arr.nbytes], dtype=torch.int64, device=dev)
        sizes = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, nbytes, group=self.group)
        sizes = sizes.cpu().tolist()
        width = max(max(sizes), 1)
        mine = torch.zeros(width, dtype=torch.uint8, device=dev)
        if arr.nbytes:
            mine[:arr.nbytes] = torch.from_numpy(np.array(arr.view(np.uint8).reshape(-1))).to(dev)   # a copy: the source may be read-only
        allb = torch.empty(self.world * width, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, mine, group=self.group)
        host = allb.cpu().numpy()
        return [host[r * width:r * width + sizes[r]].view(arr.dtype).copy() for r in range(self.world)]


class ThreadComm:
    """Same exchanges between the threads of ONE process, one thread per shard (one process driving several
    GPUs, or several handles on one GPU): ThreadComm.make(world) -> list of endpoints."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self._sh = rank, world, shared

    @staticmethod
    def make(world):
        import threading
        shared = {"barrier": threading.Barrier(world), "slots": [None] * world}
        return [ThreadComm(r, world, shared) for r in range(world)]

    def all_gather(self, obj):
        sh = self._sh
        sh["slots"][self.rank] = obj
        sh["barrier"].wait()
        out = list(sh["slots"])
        sh["barrier"].wait()
        return out

    def all_gather_array(self, arr):
        import numpy as np
        return [np.array(a, copy=True) for a in self.all_gather(np.ascontiguousarray(arr))]

    def all_reduce_sum_(self, t):
        import torch
        if t.is_cuda:
            torch.cuda.synchronize(t.device)
        parts = self.all_gather(t)
        total = parts[0].to(t.device, copy=True)
        for q in parts[1:]:
            total += q.to(t.device)
        if t.is_cuda:
            torch.cuda.synchronize(t.device)
        self._sh["barrier"].wait()            # everyone has read every table before anyone overwrites its own
        t.copy_(total)
        if t.is_cuda:
            torch.cuda.synchronize(t.device)
        self._sh["barrier"].wait()


def create_pdws_sharded(shard, comm=None):
    """create_pdws_channelized.m:60-136 over a recording whose channel matrix is time-sharded, one shard per
    rank of `comm` (rank order = time order; default: the torch.distributed world).  Every rank calls this
    with its PdwShard.  Exchanges: 3 sums of the histogram table (device memory; NCCL all-reduce), and
    typed all-gathers (lengths + padded bytes, TorchDistComm.all_gather_array) of the exit codes (M bytes), the
    edge events, the column segments of boundary pulses and the finished records.  Returns (records as _lib.Pdw in the reference's order, noise floor per
    natural channel) on every rank."""
    import ctypes as C
    import numpy as np
    from . import _lib
    comm = comm or TorchDistComm()
    rank, world = comm.rank, comm.world
    gather = comm.all_gather_array                                         # typed arrays, no pickling

    # 1. exact median of |y| per channel over the whole recording (:73): sum the shards' histograms
    for p in range(3):
        comm.all_reduce_sum_(shard.hist(p))
        shard.select(p)
    shard.thresholds()                                                     # :74-75
    # 2. FSM state on entry to this shard
    bounds = [tuple(int(v) for v in b) for b in gather(np.array([shard.row_offset, shard.nrows], dtype=np.int64))]
    codes = gather(shard.exit_state())
    entry = fold_exit_codes(codes[:rank], shard.M)
    # 3. edges (:79-96), pulses in the reference's order (same list on every rank)
    events = np.concatenate(gather(shard.detect(entry)))
    pulses = pair_events(events, shard.M, shard.bug)

    def owner(row):                                                       # rank holding 1-based row `row`
        for r, (off, n) in enumerate(bounds):
            if off < row <= off + n:
                return r
        raise ValueError(f"row {row} outside every shard")

    own = [(owner(p.toa_row), owner(p.end_row)) for p in pulses]
    mine = [i for i, (a, b) in enumerate(own) if a == b == rank]
    recs = dict(zip(mine, shard.records([pulses[i] for i in mine])))
    # 4. pulses that straddle shards: every rank contributes the rows it holds of the pulse's column(s);
    #    the rank holding the trailing edge assembles them and computes the record
    off, n = bounds[rank]
    seg_idx, seg_rows = [], []
    for i, (a, b) in enumerate(own):
        if a == b:
            continue
        p = pulses[i]
        lo, hi = max(p.toa_row, off + 1), min(p.end_row, off + n)
        if lo <= hi:
            seg = shard.column_segment((p.col, p.col_phase), lo, hi)       # complex64 [rows, 2]
            seg_idx.append((i, seg.shape[0]))
            seg_rows.append(np.ascontiguousarray(seg, dtype=np.complex64).reshape(-1))
    all_idx = gather(np.array(seg_idx, dtype=np.int64).reshape(-1))         # (pulse index, rows) pairs per rank
    all_dat = gather(np.concatenate(seg_rows) if seg_rows else np.zeros(0, dtype=np.complex64))
    all_segs = []
    for idx, dat in zip(all_idx, all_dat):
        d, pos = {}, 0
        for i, nr in idx.reshape(-1, 2):
            d[int(i)] = dat[pos:pos + 2 * int(nr)].reshape(int(nr), 2)
            pos += 2 * int(nr)
        all_segs.append(d)
    for i, (a, b) in enumerate(own):
        if a == b or b != rank:
            continue
        p = pulses[i]
        mat = np.concatenate([all_segs[r][i] for r in range(world) if i in all_segs[r]], axis=0)
        assert mat.shape[0] == p.end_row - p.toa_row + 1
        q = _lib.Pulse(p.toa_row, p.end_row, p.channel_natural, 0, 1, 0)
        recs[i] = shard.records_from_matrix(mat, p.toa_row - 1, [q])[0]
    # 5. records: (pulse index, 80 bytes of chz_pdw_t) per rank
    rsz = C.sizeof(_lib.Pdw)
    keys = sorted(recs)
    all_keys = gather(np.array(keys, dtype=np.int64))
    all_recs = gather(np.frombuffer(b"".join(recs[i] for i in keys), dtype=np.uint8) if keys else np.zeros(0, dtype=np.uint8))
    merged = {}
    for ks, blob in zip(all_keys, all_recs):
        for j, i in enumerate(ks):
            merged[int(i)] = blob[j * rsz:(j + 1) * rsz].tobytes()
    out = [_lib.Pdw.from_buffer_copy(merged[i]) for i in range(len(pulses))]
    return out, shard.noise_floor()
