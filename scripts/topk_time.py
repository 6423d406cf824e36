"""Fused scoring + mask + top-K timing (incl. the merge); random tables, realistic train masks are not needed for the
timing (the mask cursor is O(history))."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ops = bench.pkg("ops")
DEV = "cuda:0"
def timeit(fn, iters=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
gen = torch.Generator().manual_seed(1)
for (nu, ni, d, k) in [(9130, 7050, 64, 50), (16716, 18357, 64, 50), (16384, 100000, 64, 50), (9130, 7050, 64, 20),
                       (17122, 23033, 128, 50)]:
    ue = torch.randn(nu, d, generator=gen).to(DEV); ie = torch.randn(ni, d, generator=gen).to(DEV)
    users = torch.arange(nu, device=DEV)
    S = ops.choose_splits(nu, ni)
    t = timeit(lambda: ops.score_mask_topk(ue, users, ie, k))
    print(f"d={d} K={k} U={nu} I={ni} splits={S}: {t*1e3:.1f} us", flush=True)
