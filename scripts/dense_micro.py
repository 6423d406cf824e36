"""Micro-benchmark of the fused d x d dense layer (dense_small.cu): forward and backward, warm L2."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ops = bench.pkg("ops")
DEV = "cuda:0"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for M, d in [(26495, 64), (7050, 64), (53955, 64), (62420, 128)]:
    x = torch.randn(M, d, device=DEV, requires_grad=True)
    W = (torch.randn(d, d, device=DEV) * 0.1).requires_grad_(True)
    b = torch.randn(d, device=DEV, requires_grad=True)
    gy = torch.randn(M, d, device=DEV)
    for act in (None, "sigmoid"):
        y = ops.dense_act(x, W, b, act)
        t_f = timeit(lambda: ops.dense_act(x, W, b, act))
        t_b = timeit(lambda: torch.autograd.grad(y, (x, W, b), gy, retain_graph=True))
        bytes_f, bytes_b = 8 * M * d, (16 if act is None else 20) * M * d
        print(f"M={M} d={d} act={act}: fwd {t_f:.1f} us ({bytes_f / t_f / 1e3:.0f} GB/s)  "
              f"bwd {t_b:.1f} us ({bytes_b / t_b / 1e3:.0f} GB/s)", flush=True)
