"""How much freshly WRITTEN data does the B200 L2 keep for a later read?  Run under
ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum and read the per-kernel DRAM bytes:
a read kernel that follows a write of X MB reads (X - retained) MB from DRAM."""
import sys
import torch
sizes = [int(s) for s in sys.argv[1:]] or [16, 32, 48, 64, 96, 128]
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
big.zero_()                      # flush: 1 GiB of writes
torch.cuda.synchronize()
for mb in sizes:
    a = torch.empty(mb << 18, dtype=torch.float32, device="cuda")
    big.zero_()                  # flush
    torch.cuda.synchronize()
    a.fill_(1.0)                 # write X MB
    s = a.sum()                  # read X MB
    torch.cuda.synchronize()
    print(mb, float(s))
