"""Isolated timing of the low-rank table kernels against the dense sequence they replace
(L2 flushed before every launch). `--once` runs one launch of each (for ncu)."""
import importlib
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = lambda s: importlib.import_module("recommendar-systems_b200." + s)
ops, optim, lib = pkg("ops"), pkg("optim"), pkg("lib")
dev = "cuda:0"
once = "--once" in sys.argv
shapes = [(7050, 4096, 64), (7050, 384, 64), (23033, 4096, 128)] if not once else [(7050, 4096, 64)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters * 1e3


for rows, cols, d in shapes:
    X = torch.nn.Parameter(torch.randn(rows, cols, device=dev))
    W = torch.randn(d, cols, device=dev) * 0.05
    dY = torch.randn(rows, d, device=dev) * 0.1
    oa = optim.FusedAdam([X], lr=1e-3)

    def lowrank():
        X._mmrec_lowrank = ops.LowRankGrad(dY, W)
        oa.step()
        X._mmrec_lowrank = None

    g2 = torch.zeros(1, dtype=torch.float64, device=dev)

    def sumsq():
        X._mmrec_lowrank = ops.LowRankGrad(dY, W)
        optim.lowrank_sumsq(X, g2)
        X._mmrec_lowrank = None

    if once:
        lowrank(); sumsq(); torch.cuda.synchronize()
        continue
    t_lr, t_sq = timed(lowrank), timed(sumsq)
    Y = torch.nn.Parameter(torch.randn(rows, cols, device=dev))
    ob = optim.FusedAdam([Y], lr=1e-3)
    Y.grad = torch.randn(rows, cols, device=dev)
    t_adam = timed(lambda: ob.step())
    t_dx = timed(lambda: ops.gemm(dY, True, W, False, rows, cols, d))
    nbytes = 24 * rows * cols
    print(f"{rows}x{cols} d={d}: table_adam_lowrank {t_lr:.1f} us ({nbytes / t_lr / 1e3:.0f} GB/s of 24 B/elem), "
          f"lowrank_sumsq {t_sq:.1f} us | dense: adam {t_adam:.1f} us + dX gemm {t_dx:.1f} us")
