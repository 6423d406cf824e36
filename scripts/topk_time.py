import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ops = bench.pkg("ops")
DEV = "cuda:0"
def timeit(fn, iters=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
gen = torch.Generator().manual_seed(1)
for (nu, ni, S, d) in [(9130, 7050, 2, 64), (16384, 100000, 1, 64), (17122, 23033, 1, 128), (16384, 100000, 1, 128)]:
    ue = torch.randn(nu, d, generator=gen).to(DEV); ie = torch.randn(ni, d, generator=gen).to(DEV)
    users = torch.arange(nu, device=DEV)
    t = timeit(lambda: ops.score_mask_topk(ue, users, ie, 50, n_splits=S))
    print(f"dbg={os.environ.get('MMREC_TOPK_DEBUG')} atmem={os.environ.get('MMREC_TOPK_ATMEM')} d={d} U={nu} I={ni} S={S}: {t*1e3:.1f} us", flush=True)
