"""Where the Python mirror's time goes in pdws_ptr on a configs[4] file (same y, 200 calls each way)."""
import ctypes as C, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import sdr_channelizer_b200 as pkg
from sdr_channelizer_b200 import _lib
from tests import synth
M, P, fs = 256, 16, 56e6
n = 5_600_000 // M * M; rows = n // M
iq, bw, _ = synth.pulsed_int16(n, M=M, seed=100, fs=fs)
d_in = torch.from_numpy(iq).cuda()
y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
ch = pkg.Channelizer(M, taps=pkg.design_prototype(M, P))
st = torch.cuda.current_stream(); ch.set_stream(st.cuda_stream)
ch.process_ptr(d_in.data_ptr(), n, bw, y.data_ptr(), rows); torch.cuda.synchronize()
L = _lib.lib(); prm = pkg.PdwParams(15.0, 0.9999, 0.0, fs, 0.0, 0, 0, 0.0)
def t(fn, k=200):
    fn(); t0 = time.perf_counter()
    for _ in range(k): fn()
    return (time.perf_counter() - t0) / k * 1e3
cnt = C.c_uint64(0); arr = (pkg.Pdw * 4096)()
out = {
 "python_pdws_ptr_ms": t(lambda: ch.pdws_ptr(y.data_ptr(), rows, fs)),
 "ctypes_cap0_ms": t(lambda: L.chz_pdws_dev(ch.handle, C.byref(prm), C.c_void_p(y.data_ptr()), rows, None, 0, C.byref(cnt))),
 "ctypes_with_records_ms": t(lambda: L.chz_pdws_dev(ch.handle, C.byref(prm), C.c_void_p(y.data_ptr()), rows, C.cast(arr, C.c_void_p), 4096, C.byref(cnt))),
 "alloc_array_ms": t(lambda: (pkg.Pdw * 2048)()),
 "pdws": int(cnt.value)}
print(json.dumps(out))
