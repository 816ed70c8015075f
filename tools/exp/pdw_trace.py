"""configs[4] PDW stage on one handle: per-file host wall clock of chz_pdws_dev through the Python mirror, through
ctypes alone, and the GPU time of the same call between two CUDA events."""
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
if os.environ.get("EVENT_PATH"):
    ch.set_option(pkg.CHZ_OPT_PDW_EVENT_PATH, 1)
st = torch.cuda.current_stream(); ch.set_stream(st.cuda_stream)
files = int(sys.argv[1]) if len(sys.argv) > 1 else 6
tot = tot_c = tot_gpu = 0.0; npdw = 0
prm = pkg.PdwParams(15.0, 0.9999, 0.0, fs, 0.0, 0, 0, 0.0)
for i in range(-2, files):
    iq, bw, _ = synth.pulsed_int16(n, M=M, seed=100 + max(i, 0), fs=fs)
    d_in = torch.from_numpy(iq).cuda(); torch.cuda.synchronize()
    ch.reset(); ch.process_ptr(d_in.data_ptr(), n, bw, y.data_ptr(), rows); torch.cuda.synchronize()
    t0 = time.perf_counter()
    recs, _ = ch.pdws_ptr(y.data_ptr(), rows, fs)
    dt = time.perf_counter() - t0
    ch.reset(); ch.process_ptr(d_in.data_ptr(), n, bw, y.data_ptr(), rows); torch.cuda.synchronize()
    cnt = C.c_uint64(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(st)
    _lib.lib().chz_pdws_dev(ch.handle, C.byref(prm), C.c_void_p(y.data_ptr()), rows, None, 0, C.byref(cnt))
    e1.record(st); dtc = time.perf_counter() - t0
    torch.cuda.synchronize()
    if i >= 0:
        tot += dt; tot_c += dtc; tot_gpu += e0.elapsed_time(e1) * 1e-3; npdw += len(recs)
print(json.dumps({"pdw_ms_per_file_python": tot / files * 1e3, "pdw_ms_per_file_c_abi": tot_c / files * 1e3,
                  "pdw_ms_per_file_gpu_events": tot_gpu / files * 1e3, "pdws": npdw, "files": files,
                  "event_path": bool(os.environ.get("EVENT_PATH"))}))
