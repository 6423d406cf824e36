"""Preference-module forward: tcgen05 kernel (MMREC_SIDE_TC=1) against the mma.sync kernel (=0), CUDA events,
training forward (7 saved tensors) and inference forward, L2 warm (inside a step the inputs were just written)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ops = bench.pkg("ops")
DEV = "cuda:0"
def timeit(fn, iters=50):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
torch.manual_seed(0)
for n in [int(a) for a in sys.argv[1:]] or [26495, 54738, 85268]:
    d = 64
    mk = lambda bias: torch.nn.Linear(d, d, bias=bias).to(DEV)
    layers = [mk(True), mk(False), mk(True), mk(False), mk(True), mk(True), mk(True)]
    ins = [torch.randn(n, d, device=DEV) for _ in range(4)]
    cnt = torch.tensor([3.0], dtype=torch.float64, device=DEV)
    for keep in (True, False):
        for tc in ("0", "1"):
            os.environ["MMREC_SIDE_TC"] = tc
            x = [t.clone().requires_grad_(keep) for t in ins]
            def run():
                with torch.set_grad_enabled(keep):
                    return ops.smore_side(*x, layers, None, (0.5, 7, cnt))
            t = timeit(run)
            print(f"n={n} saved={keep} MMREC_SIDE_TC={tc}: {t*1e3:.1f} us", flush=True)
# training forward + backward on the mma.sync kernels (the backward has no tcgen05 version)
os.environ["MMREC_SIDE_TC"] = "0"
for n in [26495]:
    x = [t.clone().requires_grad_(True) for t in ins] if ins[0].shape[0] == n else None
    if x is None:
        ins = [torch.randn(n, 64, device=DEV) for _ in range(4)]
        x = [t.clone().requires_grad_(True) for t in ins]
    ga, gs = torch.randn(n, 64, device=DEV), torch.randn(n, 64, device=DEV)
    def run_fb():
        a, s = ops.smore_side(*x, layers, None, (0.5, 7, cnt))
        torch.autograd.backward([a, s], [ga, gs])
    print(f"n={n} fwd+bwd (mma.sync): {timeit(run_fb)*1e3:.1f} us", flush=True)
