"""What does a trivial kernel with the fused channelizer's traffic mix (4 B read + 8 B written per sample) reach?
Times K1 alone (chz_unpack_dev: int16 pair -> float2) at the headline size next to torch copies, so that the fused
kernel's fraction of the copy peak can be read against the best any 1:2 read:write kernel does on this board."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sdr_channelizer_b200 as pkg

n = int(os.environ.get("N", "614400000"))
x = torch.randint(-2048, 2048, (n, 2), dtype=torch.int16, device="cuda")
y = torch.empty((n,), dtype=torch.complex64, device="cuda")
st = torch.cuda.current_stream()

def timed(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
ms = timed(lambda: pkg.unpack_ptr(x.data_ptr(), n, 12, y.data_ptr(), st.cuda_stream))
out["k_unpack_4B_in_8B_out"] = {"ms": round(ms, 4), "GBps": round(12 * n / ms / 1e6, 1)}
a = torch.empty(n * 3 // 2, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)      # same 12 B per sample in total, 1:1
ms = timed(lambda: b.copy_(a))
out["torch_copy_6B_in_6B_out"] = {"ms": round(ms, 4), "GBps": round(12 * n / ms / 1e6, 1)}
ms = timed(lambda: y.zero_())
out["memset_8B_out"] = {"ms": round(ms, 4), "GBps": round(8 * n / ms / 1e6, 1)}
yv = torch.view_as_real(y)
ms = timed(lambda: yv.sum())
out["read_only_8B_in"] = {"ms": round(ms, 4), "GBps": round(8 * n / ms / 1e6, 1)}
print(json.dumps(out))
