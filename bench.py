#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json configs[1]):

    64-channel critically sampled channelizer, 1024-tap prototype, on a 12-bit bladeRF-format
    recording at 61.44 MS/s, 10 s long (614.4 M complex int16 samples = 2.46 GB in, 4.92 GB out).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of unpack -> polyphase FIR -> FFT over one recording.  N GPUs: the recording is N
times longer and is sharded along time, rank r taking 10 s plus a (taps-1)-sample halo — no
collective on the data path (weak scaling).  Prints ONE JSON line on rank 0.  Besides the contract's
keys the line carries
  roofline      dominant kernel against the measured HBM peak (+ frac_sustained over 200 back-to-back launches)
  cpu_baseline  the oracle port on the host cores, bounded sample
  e2e           the same metric through chz_process() with pinned HOST buffers (+ the plain-copy ceilings of
                the same buffers, so the record shows how close the pipeline is to the host's limit)
  e2e_pdw       recording in host memory -> PDWs (chz_process(out = NULL) + chz_pdws): the reference script's use
  parity        full-size spot check: random rows of the timed output against the oracle run on just the samples
                those rows touch (headline; at N = 1 also configs[3], including rows past 2^32 / M)
  others        (N = 1) every other BASELINE config, device resident, a few steps each
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M, TAPS_PER_BAND, OVERSAMPLE, BIT_WIDTH = 64, 16, 1, 12
FS = 61_440_000
SECONDS = 10
N_SAMPLES = FS * SECONDS                      # per GPU
BYTES_PER_SAMPLE_ALGO = 4 + 8 * OVERSAMPLE    # int16 pair in, fp32 complex out
METRIC = "input complex MS/s channelized"
WORKLOAD = ("configs[1]: 64-channel critically sampled channelizer, 1024-tap prototype, 12-bit bladeRF-format "
            "recording at 61.44 MS/s, 10 s per GPU")
# the other BASELINE.json configs (device-resident side measurements at N = 1): M, taps/band, oversample, bits, samples
OTHERS = {
    "configs[0]": (8, 8, 1, 8, 1_000_000, "8-channel critically sampled, 64 taps, 1M-sample 8-bit file"),
    "configs[2]": (1024, 16, 2, 16, 560_000_000, "1024-channel 2x oversampled, 16384 taps, 56 MS/s x 10 s, 16-bit container"),
    "configs[3]": (4096, 16, 1, 12, 3_686_400_000, "4096-channel, 65536 taps, 61.44 MS/s x 60 s (one GPU's view of the whole recording)"),
}


def config_dict(world, n_own):
    ntaps = M * TAPS_PER_BAND
    return {"workload": WORKLOAD, "channels": M, "taps": ntaps, "oversample": OVERSAMPLE, "bit_width": BIT_WIDTH,
            "samples_per_gpu": n_own,
            "halo_samples_per_shard": int(ntaps - 1 + (2 * M - (ntaps - 1) % (2 * M)) % (2 * M)) if world > 1 else 0,
            "parallelism": f"time-sharded x{world}, no collective",
            "l2": "inputs (2.46 GB) and outputs (4.92 GB) per step exceed the 126 MB L2; no flush needed"}


def _peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _traffic_from_profiles():
    try:
        with open(os.path.join(ROOT, "profiles", "fused_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch_full_workload")
    except Exception:
        return None


def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this process (and therefore the pages of every pinned buffer it allocates afterwards: first touch)
    to the NUMA node the GPU hangs off.  Eight ranks that all stage through one socket's memory were the reason
    the end-to-end figure did not scale in round 1."""
    info = {"node": None, "cpus": None}
    bus = None
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:
        bus = None
    try:
        if bus is None:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
            bus = out.splitlines()[0].strip() if out else None
        if not bus:
            return info
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as e:                         # no sysfs / not permitted: run unbound and say so
        info["error"] = repr(e)[:120]
    return info


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.2] or [l for (_, l) in self.lines]
        for line in rows:
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except Exception:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_input(torch, n, seed, device):
    """configs[1] content: 8 tones + AWGN (sigma 0.05 FS) in 12-bit Q11, clipped to [-2048, 2047]."""
    out = torch.empty((n, 2), dtype=torch.int16, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    freqs = [(-27.3 + 7.1 * i) / M for i in range(8)]
    chunk = 1 << 24
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        t = torch.arange(s, e, device=device, dtype=torch.float64)
        re = torch.randn(e - s, device=device, generator=g) * 0.05
        im = torch.randn(e - s, device=device, generator=g) * 0.05
        for f in freqs:
            ph = (2.0 * torch.pi) * torch.frac(t * f)
            re += 0.08 * torch.cos(ph).float()
            im += 0.08 * torch.sin(ph).float()
        out[s:e, 0] = torch.clamp(torch.round(re * 2048.0), -2048, 2047).to(torch.int16)
        out[s:e, 1] = torch.clamp(torch.round(im * 2048.0), -2048, 2047).to(torch.int16)
    return out


def cpu_baseline(taps, target_seconds=12.0):
    """The double-precision OpenMP oracle (kind 'port': MATLAB's dsp.Channelizer cannot run here) timed on
    this box's host cores on a bounded prefix of the same workload."""
    import numpy as np
    from oracle import pyoracle as orc
    from tests import synth
    orc.lib().orc_set_num_threads(_host_threads())             # torchrun exports OMP_NUM_THREADS=1
    h = taps.astype(np.float64)
    probe_n = M * 32768
    iq, bw = synth.tones_int16_q11(probe_n, M, seed=2)
    orc.channelize_raw(iq, bw, M, h, OVERSAMPLE)            # warm-up (threads, page faults)
    t0 = time.perf_counter(); orc.channelize_raw(iq, bw, M, h, OVERSAMPLE); dt = time.perf_counter() - t0
    rate = probe_n / dt
    n = int(min(FS, max(probe_n, rate * target_seconds / 3))) // M * M      # at most a 1 s prefix
    reps = max(1, int(np.ceil(n / probe_n)))
    big = np.tile(iq, (reps, 1))[:n]
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); orc.channelize_raw(big, bw, M, h, OVERSAMPLE); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": n / best / 1e6, "unit": "MS/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"first {n} samples ({n / FS:.3f} s) of the recording, best of 3, oracle/chz_oracle.cpp "
                      f"(double precision, OpenMP)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Its arithmetic is MATLAB's
    closed-source dsp.Channelizer (not runnable here), so this arm times the oracle port on the host
    cores.  Same config, metric and unit as the B200 arm; every one of the --steps steps is a bounded sample
    (a prefix of the recording, sized so that the whole run ends within a few minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import pyoracle as orc
    from tests import synth
    orc.lib().orc_set_num_threads(_host_threads())             # torchrun exports OMP_NUM_THREADS=1
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    taps = orc.design_prototype(M, TAPS_PER_BAND)
    n_own = int(FS * args.seconds) // M * M
    base, bw = synth.tones_int16_q11(M * 32768, M, seed=2)
    # probe the host's rate, then size the per-step prefix so that warm-up + K steps stay under ~150 s
    t0 = time.perf_counter(); orc.channelize_raw(base, bw, M, taps, OVERSAMPLE); orc.channelize_raw(base, bw, M, taps, OVERSAMPLE)
    rate = 2 * len(base) / (time.perf_counter() - t0)
    total_steps = max(1, args.steps) + max(1, min(args.warmup, 3))
    n = int(min(FS, max(len(base), rate * 150.0 / total_steps))) // M * M           # at most a 1 s prefix per step
    big = np.tile(base, (n // len(base) + 1, 1))[:n]
    for _ in range(max(1, min(args.warmup, 3))):
        orc.channelize_raw(big, bw, M, taps, OVERSAMPLE)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.channelize_raw(big, bw, M, taps, OVERSAMPLE)
    dt = (time.perf_counter() - t0) / steps
    val = n / dt / 1e6
    sample = (f"each step = the first {n} samples ({n / FS:.3f} s) of the recording (bounded CPU sample); timed: "
              f"oracle/chz_oracle.cpp port of the path, double precision, OpenMP on all host threads (the reference's own "
              f"arithmetic is MATLAB's closed-source dsp.Channelizer; no MATLAB/Octave here)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "MS/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world, n_own),
            "cpu_baseline": {"value": val, "unit": "MS/s", "cores": orc.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def spot_parity(torch, orc, x, y, m_, p_, os_, bw, taps, count=32, seed=0):
    """Random rows of a full-size device-resident result against the double-precision oracle evaluated on just
    the samples those rows touch (the window keeps the row's frame and rotation phase; x[n < 0] = 0)."""
    import numpy as np
    from tests import synth
    d_, l_ = m_ // os_, m_ * p_
    rows = y.shape[0]
    rng = np.random.default_rng(seed)
    picks = [0, 1, os_ * p_ - 1, rows - 2, rows - 1] + [int(v) for v in rng.integers(p_, rows, count - 5)]
    # rows on both sides of the places where 32-bit index arithmetic would wrap: element index 2^31 and 2^32,
    # byte offset 2^32 (configs[3]: 900 000 rows x 4096 channels = 3.69e9 elements, 29.5 GB)
    k = 5
    for edge in ((1 << 31) // m_, (1 << 32) // m_, (1 << 32) // (8 * m_)):
        if p_ < edge < rows - 2 and k + 2 <= len(picks):
            picks[k:k + 2] = [edge - 1, edge + 1]
            k += 2
    worst, h = 0.0, taps.astype(np.float64)
    for m in picks:
        lo = max(0, m * d_ - (l_ - 1))
        seg = x[lo:m * d_ + 1].cpu().numpy()
        seg_full = np.concatenate([np.zeros((l_ - len(seg), 2), seg.dtype), seg])
        idx = l_ + (m * d_) % m_
        win = np.concatenate([np.zeros((idx - (l_ - 1), 2), seg.dtype), seg_full, np.zeros((d_ - 1, 2), seg.dtype)])
        ref = orc.channelize_raw(win, bw, m_, h, os_, row0=idx // d_, nrows=1)[0]
        got = y[m].cpu().numpy()
        worst = max(worst, float(synth.rel_rms(got, ref)))
    return {"rows_checked": len(picks), "max_rel_rms": worst, "tolerance": 1e-5, "rows_total": int(rows),
            "max_row_index_checked": int(max(picks)), "ok": bool(worst <= 1e-5)}


def run_other(torch, pkg, orc, name, peak, steps=3):
    """Device-resident throughput of one of the other BASELINE configs (same method as the headline), plus the
    full-size parity spot check for configs[3]."""
    m_, p_, os_, bw, n, note = OTHERS[name]
    n = n // m_ * m_
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(7)
    lim = 2 ** (bw - 1)
    x = torch.empty((n, 2), dtype=torch.int8 if bw <= 8 else torch.int16, device=dev)
    chunk = 1 << 28
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x[s:e] = torch.randint(-lim, lim, (e - s, 2), dtype=x.dtype, device=dev, generator=g)
    rows = n // (m_ // os_)
    y = torch.empty((rows, m_), dtype=torch.complex64, device=dev)
    taps = pkg.design_prototype(m_, p_)
    ch = pkg.Channelizer(m_, taps=taps, OversamplingRatio=os_)
    st = torch.cuda.current_stream()
    ch.set_stream(st.cuda_stream)
    for _ in range(2):
        ch.reset(); ch.process_ptr(x.data_ptr(), n, bw, y.data_ptr(), rows)
    l0 = ch.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        ch.reset(); ch.process_ptr(x.data_ptr(), n, bw, y.data_ptr(), rows)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    bps = (2 if bw <= 8 else 4) + 8 * os_
    out = {"config": name, "note": note, "channels": m_, "taps": m_ * p_, "oversample": os_, "bit_width": bw, "samples": n,
           "steps": steps, "ms": ms, "MS_per_s": n / (ms * 1e-3) / 1e6, "algorithmic_bytes_per_sample": bps,
           "frac": bps * n / (ms * 1e-3) / 1e9 / peak, "launches_per_step": (ch.kernel_launches - l0) / steps}
    if name in ("configs[3]", "configs[2]"):
        out["parity"] = spot_parity(torch, orc, x, y, m_, p_, os_, bw, taps, count=32 if name == "configs[3]" else 12)
    ch.close()
    del x, y
    torch.cuda.empty_cache()
    return out


def run_cfg4_pdw(torch, pkg, files=8):
    """configs[4]: pulsed files (100 ms @ 56 MS/s, int16) -> 256 channels -> PDWs on one handle; per-file host time
    of the two stages (the PDW stage ends with its records on the host).  The files are synthesised and uploaded
    first and one untimed pass warms the handle up, so that the timed passes see a busy GPU (as a batch job does)
    rather than one that idled through a second of CPU-side synthesis per file."""
    from tests import synth
    m_, p_, fs = 256, 16, 56e6
    n = 5_600_000 // m_ * m_
    rows = n // m_
    y = torch.empty((rows, m_), dtype=torch.complex64, device="cuda")
    ch = pkg.Channelizer(m_, taps=pkg.design_prototype(m_, p_))
    st = torch.cuda.current_stream()
    ch.set_stream(st.cuda_stream)
    d_in, bw = [], 16
    for i in range(files):
        iq, bw, _ = synth.pulsed_int16(n, M=m_, seed=100 + i, fs=fs)
        d_in.append(torch.from_numpy(iq).cuda())
    torch.cuda.synchronize()
    best = None
    for rep in range(-1, 3):                    # pass -1 is an untimed warm-up; best of three timed passes
        tot_chan = tot_pdw = 0.0
        npdw = 0
        for i in range(files):
            ch.reset()
            t0 = time.perf_counter()
            ch.process_ptr(d_in[i].data_ptr(), n, bw, y.data_ptr(), rows); torch.cuda.synchronize()
            t1 = time.perf_counter()
            recs, _ = ch.pdws_ptr(y.data_ptr(), rows, fs)
            t2 = time.perf_counter()
            tot_chan += t1 - t0; tot_pdw += t2 - t1; npdw += len(recs)
        if rep >= 0 and (best is None or tot_chan + tot_pdw < best[0] + best[1]):
            best = (tot_chan, tot_pdw, npdw)
    tot_chan, tot_pdw, npdw = best
    ch.close()
    return {"config": "configs[4]", "note": "8 pulsed files (100 ms @ 56 MS/s, int16) -> 256 channels -> PDWs, one handle, "
            "device-resident input, host-timed per file (files uploaded first, one warm-up pass, best of 3 passes)",
            "files": files, "samples_per_file": n, "pdws": npdw,
            "chan_ms_per_file": tot_chan / files * 1e3, "pdw_ms_per_file": tot_pdw / files * 1e3,
            "MS_per_s": files * n / (tot_chan + tot_pdw) / 1e6, "files_per_s": files / (tot_chan + tot_pdw),
            "pdws_per_s": npdw / (tot_chan + tot_pdw)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--seconds", type=float, default=float(SECONDS), help="recording length per GPU (default: the config's 10 s)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the side measurements of the other BASELINE configs (N = 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import sdr_channelizer_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: sdr_channelizer_b200 has no CPU path")
    numa = bind_to_gpu_numa_node(torch, local_rank)       # before any pinned allocation and before CUDA spawns its threads
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    n_own = int(FS * args.seconds) // M * M
    ntaps = M * TAPS_PER_BAND

    # time shard of a world*seconds recording: own samples plus the (taps-1)-sample halo, frame aligned
    shard = pkg.plan_time_shards(n_own * world, M, ntaps, OVERSAMPLE, world)[rank]
    x = make_input(torch, shard.samples, 2 + rank, dev)           # [halo + own, 2] int16, device resident
    rows_total = shard.samples // (M // OVERSAMPLE)
    y = torch.empty((rows_total, M), dtype=torch.complex64, device=dev)
    taps = pkg.design_prototype(M, TAPS_PER_BAND)
    ch = pkg.Channelizer(M, taps=taps, OversamplingRatio=OVERSAMPLE)
    if os.environ.get("CHZ_BENCH_PATH"):                          # kernel A/B experiments only
        ch.set_option(pkg.CHZ_OPT_FORCE_PATH, int(os.environ["CHZ_BENCH_PATH"]))
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)        # the kernels AND the timing events live on this stream
    torch.cuda.set_stream(stream)
    ch.set_stream(stream.cuda_stream)

    def step():
        ch.reset()
        return ch.process_ptr(x.data_ptr(), shard.samples, BIT_WIDTH, y.data_ptr(), rows_total)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(warmup):
        step()
    barrier()
    launches0 = ch.kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_wall0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        rows = step()
        ev[i + 1].record(stream)
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    assert rows == rows_total
    per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    launches = ch.kernel_launches - launches0
    ms_per_step = max_over_ranks(total_ms) / args.steps
    value = (n_own * world) / (ms_per_step * 1e-3) / 1e6      # owned samples of all ranks / max time

    # What a do-nothing kernel with the SAME traffic reaches on this board: K1 alone (int16 pair -> float2: 4 B read,
    # 8 B written per sample) over the same buffers.  The copy peak in MEASURED_PEAKS.json is a 1:1 read:write mix;
    # this path writes twice what it reads.  K1 (two samples per thread, 8-byte loads, 16-byte stores) reaches ~95 %
    # of the copy peak with that mix (tools/ubench/mixbw.cu has the variants), so the gap between the fused kernel
    # and K1 is what the FIR and the FFT cost on top of the traffic.
    # Measured right after the timed region, before the 200-launch sustained run heats the board.
    same_traffic = None
    if BIT_WIDTH > 8 and OVERSAMPLE == 1:
        nev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        for _ in range(2):
            pkg.unpack_ptr(x.data_ptr(), shard.samples, BIT_WIDTH, y.data_ptr(), stream.cuda_stream)
        nev[0].record(stream)
        for i in range(10):
            pkg.unpack_ptr(x.data_ptr(), shard.samples, BIT_WIDTH, y.data_ptr(), stream.cuda_stream)
            nev[i + 1].record(stream)
        barrier()
        same_traffic = statistics.mean(nev[i].elapsed_time(nev[i + 1]) for i in range(10))

    # sustained figure: 200 more launches back to back (the board's power management lowers the SM clock after
    # ~50 ms of this kernel; the timed region above is whatever --steps asked for)
    sus_n = 200
    sev = [torch.cuda.Event(enable_timing=True) for _ in range(sus_n + 1)]
    sev[0].record(stream)
    for i in range(sus_n):
        step()
        sev[i + 1].record(stream)
    barrier()
    sus_ms = statistics.mean(sev[i].elapsed_time(sev[i + 1]) for i in range(sus_n // 2, sus_n))

    # roofline of the dominant kernel (fused unpack+FIR+FFT: one launch per step), rank-0 numbers
    peak, peak_src = _peak_hbm()
    kern_ms = statistics.mean(per_step_ms)
    algo_bytes = shard.samples * BYTES_PER_SAMPLE_ALGO
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _traffic_from_profiles(),
                "traffic_source": "profiles/fused_traffic.json: dram__bytes_read + dram__bytes_write of this kernel at this size from "
                                  "an ncu --set full capture (profiles/r01h_fused64_full_size_ncu.txt), not measured in this run",
                "kernel": "k_chan_fused<64,16,int16>",
                "algorithmic_bytes_per_launch": algo_bytes,
                "ms_per_launch": kern_ms, "ms_per_launch_min": min(per_step_ms),
                "ms_per_launch_median": statistics.median(per_step_ms), "peak_source": peak_src,
                "frac_sustained": algo_bytes / (sus_ms * 1e-3) / 1e9 / peak,
                "sustained": f"mean of launches 100..199 of {sus_n} back-to-back launches after the timed region: {sus_ms:.4f} ms"}
    if same_traffic:
        roofline["same_traffic_noop"] = {"kernel": "k_unpack (K1 alone: 4 B read + 8 B written per sample, no FIR, no FFT)",
                                         "ms_per_launch": same_traffic, "GBps": algo_bytes / (same_traffic * 1e-3) / 1e9,
                                         "frac_of_peak": algo_bytes / (same_traffic * 1e-3) / 1e9 / peak,
                                         "this_kernel_vs_noop": same_traffic / kern_ms}
    if os.environ.get("CHZ_BENCH_DUMP"):      # per-step series (power-cap / clock drift diagnosis)
        k = max(1, len(per_step_ms) // 10)
        print("per-step ms, means of consecutive tenths:", [round(statistics.mean(per_step_ms[i:i + k]), 4) for i in range(0, len(per_step_ms), k)],
              file=sys.stderr, flush=True)

    # full-size parity of the timed output (outside the timed region): rank 0 checks its shard
    parity = None
    orc = None
    if rank == 0 and not args.no_cpu:
        from oracle import pyoracle as orc
        orc.lib().orc_set_num_threads(_host_threads())
        step(); torch.cuda.synchronize()
        parity = {"configs[1]": spot_parity(torch, orc, x, y, M, TAPS_PER_BAND, OVERSAMPLE, BIT_WIDTH, taps, count=32)}

    # end to end through the C ABI with HOST buffers: pinned input -> H2D -> kernels -> D2H -> pinned output
    e2e = e2e_pdw = None
    if not args.no_e2e:
        h_in = torch.empty((shard.samples, 2), dtype=torch.int16, pin_memory=True)
        h_in.copy_(x)
        h_out = torch.empty((rows_total, M), dtype=torch.complex64, pin_memory=True)
        h_out.zero_()                                           # first touch on this rank's NUMA node
        ch.retain(False)
        e_steps = max(2, min(5, args.steps))

        def e2e_step():
            ch.reset()
            return ch.process_ptr(h_in.data_ptr(), shard.samples, BIT_WIDTH, h_out.data_ptr(), rows_total, device=False)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()                                          # synchronous: returns when the output is on the host
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / e_steps)
        e2e = {"value": (n_own * world) / dt / 1e6, "unit": "MS/s",
               "h2d_bytes_per_step": int(h_in.numel() * 2), "d2h_bytes_per_step": int(h_out.numel() * 8),
               "ms_per_step": dt * 1e3, "steps": e_steps, "numa": numa,
               "how": "chz_process() on pinned host buffers, chunked H2D/kernel/D2H pipeline inside the call"}
        # the host-path output equals the device-path output: WHOLE buffer, compared on the device in chunks
        step(); torch.cuda.synchronize()
        same = True
        crow = max(1, (256 << 20) // (M * 8))
        for r0 in range(0, rows_total, crow):
            r1 = min(rows_total, r0 + crow)
            same = same and bool(torch.equal(h_out[r0:r1].to(dev, non_blocking=False).view(torch.float32), y[r0:r1].view(torch.float32)))
        e2e["matches_device_path"] = same
        e2e["match_scope"] = f"all {rows_total} rows bit for bit"
        # plain-copy ceilings of the same pinned buffers: all ranks copy at the same time, as in the pipeline
        def copy_rate(fn, nbytes, reps=3):
            fn(); barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            barrier()
            return nbytes / (max_over_ranks((time.perf_counter() - t0) / reps)) / 1e9
        h2d = copy_rate(lambda: (x.copy_(h_in, non_blocking=True), torch.cuda.synchronize()), h_in.numel() * 2)
        d2h = copy_rate(lambda: (h_out.copy_(y, non_blocking=True), torch.cuda.synchronize()), h_out.numel() * 8)
        s2 = torch.cuda.Stream(device=dev)

        def both():
            with torch.cuda.stream(s2):
                x.copy_(h_in, non_blocking=True)
            h_out.copy_(y, non_blocking=True)
            torch.cuda.synchronize()
        t_both = (h_in.numel() * 2 + h_out.numel() * 8) / copy_rate(both, h_in.numel() * 2 + h_out.numel() * 8) / 1e9
        e2e["h2d_gbs"] = h2d                                    # per rank, with every rank copying
        e2e["d2h_gbs"] = d2h
        e2e["copy_ceiling_ms"] = t_both * 1e3                   # both directions at once, plain cudaMemcpyAsync
        e2e["frac_of_copy_ceiling"] = t_both / dt
        del h_out

        # recording in host memory -> PDWs: the channel matrix never leaves the GPU (create_pdws_channelized.m's use)
        ch.retain(True)
        n_pdw = None

        def pdw_step():
            ch.reset()
            ch.process_ptr(h_in.data_ptr(), shard.samples, BIT_WIDTH, 0, 0, device=False)
            recs, _ = ch.pdws(float(FS))
            return len(recs)

        pdw_step()
        barrier()
        t0 = time.perf_counter()
        p_steps = 2
        for _ in range(p_steps):
            n_pdw = pdw_step()
        barrier()
        dtp = max_over_ranks((time.perf_counter() - t0) / p_steps)
        e2e_pdw = {"value": (n_own * world) / dtp / 1e6, "unit": "MS/s", "ms_per_step": dtp * 1e3, "steps": p_steps,
                   "h2d_bytes_per_step": int(h_in.numel() * 2), "d2h_bytes_per_step": int(n_pdw * C.sizeof(pkg.Pdw)),
                   "pdws_rank0": int(n_pdw), "note": "every rank runs the whole script on its own 10 s recording (replicas); "
                   "chz_process(out = NULL, CHZ_OPT_RETAIN) + chz_pdws"}
        ch.retain(False)
        ch.reset()
        del h_in

    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline(taps, target_seconds=12.0 if world == 1 else 5.0)

    others = None
    if rank == 0 and world == 1 and not args.no_others:
        del x, y
        torch.cuda.empty_cache()
        others = []
        for name in ("configs[0]", "configs[2]", "configs[3]"):
            try:
                r = run_other(torch, pkg, orc, name, peak) if orc is not None else None
                if r and "parity" in r:
                    parity[name] = r.pop("parity")
            except Exception as e:      # one config must not hide the others
                r = {"config": name, "error": repr(e)[:200]}
            if r:
                others.append(r)
        try:
            others.append(run_cfg4_pdw(torch, pkg))
        except Exception as e:
            others.append({"config": "configs[4]", "error": repr(e)[:200]})

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(world, n_own),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_pdw": e2e_pdw, "gpu_launches": int(launches),
                "clocks": clocks, "parity": parity, "others": others}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ch.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
