"""Stress of the M = 1024 ring kernel (role-split, named barriers, TMA ring): random lengths, chunkings, bit widths,
oversampling, taps per band and input alignments.  Each case runs three times through the default path -- the
outputs must be bit-identical (a race would show as nondeterminism) -- and once through the split path
(CHZ_OPT_FORCE_PATH = 2), which must agree to rel-RMS 1e-5 with it."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import sdr_channelizer_b200 as pkg

M = 1024
rng = np.random.default_rng(int(os.environ.get("SEED", "7")))
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
bad = 0
worst = 0.0
for c in range(cases):
    P = int(rng.choice([8, 12, 16])); os_ = int(rng.choice([1, 2])); bits = int(rng.choice([8, 12, 16]))
    D = M // os_
    frames = int(rng.choice([1, 3, 8, 17, 40, 129, 700, 5000]))
    n = frames * M + int(rng.integers(0, M))                    # ragged tail: the library trims to whole frames
    off = int(rng.choice([0, 0, 1, 3, 4]))                      # element offset of the input pointer: 0 = 16-byte aligned
    nchunks = int(rng.choice([1, 1, 2, 5]))
    lim = 2 ** (bits - 1)
    dt = torch.int8 if bits <= 8 else torch.int16
    xb = torch.randint(-lim, lim, (n + 8, 2), dtype=dt, device="cuda")
    x = xb[off:off + n]
    cuts = sorted(set([0, n] + [int(v) // D * D for v in rng.integers(0, n + 1, nchunks - 1)]))
    rows_cap = n // D + 8
    outs = []
    for run in range(4):
        ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P), OversamplingRatio=os_)
        if run == 3:
            ch.set_option(pkg.CHZ_OPT_FORCE_PATH, 2)
        y = torch.zeros((rows_cap, M), dtype=torch.complex64, device="cuda")
        r = 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b > a:
                r += ch.process_ptr(x[a:].data_ptr(), b - a, bits, y[r:].data_ptr(), rows_cap - r)
        torch.cuda.synchronize()
        outs.append((y[:r].clone(), r))
        ch.close()
    same = all(o[1] == outs[0][1] and torch.equal(o[0].view(torch.float32), outs[0][0].view(torch.float32)) for o in outs[1:3])
    rel = 0.0
    if outs[0][1]:
        d = (outs[0][0] - outs[3][0]).abs().pow(2).sum().sqrt().item()
        rel = d / max(outs[3][0].abs().pow(2).sum().sqrt().item(), 1e-30)
    ok = same and outs[3][1] == outs[0][1] and rel <= 1e-5
    worst = max(worst, rel)
    if not ok:
        bad += 1
        print(json.dumps({"case": c, "P": P, "os": os_, "bits": bits, "n": n, "off": off, "cuts": cuts, "rows": [o[1] for o in outs],
                          "repeatable": same, "rel_rms_vs_split": rel}), flush=True)
print(json.dumps({"cases": cases, "failed": bad, "worst_rel_rms_vs_split_path": worst}))
sys.exit(1 if bad else 0)
