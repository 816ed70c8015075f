"""Python mirror against bare ctypes on the same channel matrix: alternating calls, per-call host time."""
import ctypes as C, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg
from sdr_channelizer_b200 import _lib
from tests import synth
M, P, fs = 256, 16, 56e6
n = 5_600_000 // M * M
rows = n // M
y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
st = torch.cuda.current_stream(); ch.set_stream(st.cuda_stream)
iq, bw, _ = synth.pulsed_int16(n, M=M, seed=100, fs=fs)
d_in = torch.from_numpy(iq).cuda(); torch.cuda.synchronize()
ch.reset(); ch.process_ptr(d_in.data_ptr(), n, bw, y.data_ptr(), rows); torch.cuda.synchronize()
prm = pkg.PdwParams(15.0, 0.9999, 0.0, fs, 0.0, 0, 0, 0.0)
L = _lib.lib()
cnt = C.c_uint64(0)
out = (_lib.Pdw * 4096)()
res = {"py": [], "c_null": [], "c_out": []}
for rep in range(12):
    t0 = time.perf_counter(); recs, _ = ch.pdws_ptr(y.data_ptr(), rows, fs); res["py"].append(time.perf_counter() - t0)
    t0 = time.perf_counter(); L.chz_pdws_dev(ch.handle, C.byref(prm), C.c_void_p(y.data_ptr()), rows, None, 0, C.byref(cnt)); res["c_null"].append(time.perf_counter() - t0)
    t0 = time.perf_counter(); L.chz_pdws_dev(ch.handle, C.byref(prm), C.c_void_p(y.data_ptr()), rows, C.cast(out, C.c_void_p), 4096, C.byref(cnt)); res["c_out"].append(time.perf_counter() - t0)
print(json.dumps({k: [round(v * 1e6, 1) for v in vs[2:]] for k, vs in res.items()}), "records", len(recs))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): ch.pdws_ptr(y.data_ptr(), rows, fs)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
