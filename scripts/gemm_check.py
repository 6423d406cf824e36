"""tcgen05 3xTF32 GEMM (gemm_tc05.cu) vs float64, all three operand layouts; timing per layout.
MMREC_GEMM_TC=0 runs the mma.sync kernel for the A/B comparison."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ops = bench.pkg("ops")
DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


def timeit(fn, iters=10):
    """Device time per call: `iters` calls captured in one CUDA graph (no Python / launch gaps)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


ok = True
gen = torch.Generator().manual_seed(5)
for I, d, F in [(7050, 64, 4096), (7050, 64, 384), (18357, 64, 4096), (23033, 128, 4096), (7050, 32, 4096),
                (1030, 64, 260)]:
    x = torch.randn(I, F, generator=gen).to(DEV)
    W = (torch.randn(d, F, generator=gen) * 0.05).to(DEV)
    b = torch.randn(d, generator=gen).to(DEV)
    dy = torch.randn(I, d, generator=gen).to(DEV)
    y = ops.gemm(x, True, W, True, I, d, F, b)                 # forward
    dW = ops.gemm(dy, False, x, False, d, F, I)                # dW = dy^T x
    dx = ops.gemm(dy, True, W, False, I, F, d)                 # dx = dy W
    torch.cuda.synchronize()
    e = (rel(y, x.double() @ W.double().T + b.double()), rel(dW, dy.double().T @ x.double()),
         rel(dx, dy.double() @ W.double()))
    t = (timeit(lambda: ops.gemm(x, True, W, True, I, d, F, b)), timeit(lambda: ops.gemm(dy, False, x, False, d, F, I)),
         timeit(lambda: ops.gemm(dy, True, W, False, I, F, d)))
    mb = 4 * I * F / 1e6
    good = all(v < 2e-6 for v in e)
    ok &= good
    print(f"I={I} d={d} F={F}: rel err fwd {e[0]:.2e} dW {e[1]:.2e} dx {e[2]:.2e} | us fwd {t[0]:.1f} dW {t[1]:.1f} "
          f"dx {t[2]:.1f} | table {mb:.0f} MB -> {mb / t[0] * 1e3:.0f} / {mb / t[1] * 1e3:.0f} / {mb / t[2] * 1e3:.0f} GB/s"
          f" {'OK' if good else 'MISMATCH'}", flush=True)
print("ALL OK" if ok else "FAILED")
