"""Two launches of the fused scoring kernel at Baby evaluation size: the command ncu wraps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ops = bench.pkg("ops")
d = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nu, ni = (9130, 7050) if len(sys.argv) < 4 else (int(sys.argv[2]), int(sys.argv[3]))
gen = torch.Generator().manual_seed(1)
ue, ie = torch.randn(nu, d, generator=gen).cuda(), torch.randn(ni, d, generator=gen).cuda()
users = torch.arange(nu, device="cuda")
for _ in range(2):
    ops.score_mask_topk(ue, users, ie, 50)
torch.cuda.synchronize()
print("done")
