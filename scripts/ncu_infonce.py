"""Two forward+backward InfoNCE pair calls at SMORE/Baby batch size: the command ncu wraps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ops = bench.pkg("ops")
gen = torch.Generator().manual_seed(3)
U, I, d, B = 19445, 7050, 64, 2048
side = torch.randn(U + I, d, generator=gen).cuda().requires_grad_(True)
content = torch.randn(U + I, d, generator=gen).cuda().requires_grad_(True)
bu = torch.randint(0, U, (B,), generator=gen).cuda()
bi = torch.randint(0, I, (B,), generator=gen).cuda()
for _ in range(2):
    l = ops.infonce_pair(side, content, U, bu, bi, 0.2)
    l.backward()
torch.cuda.synchronize()
print("done")
