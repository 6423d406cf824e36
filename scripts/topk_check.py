"""tcgen05 score+mask+top-K vs the CUDA-core kernel and a float64 stable sort; timings of both."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ops = bench.pkg("ops")
DEV = "cuda:0"


def case(n_users, n_items, d, k, splits, with_mask=True, seed=0):
    gen = torch.Generator().manual_seed(seed)
    ue = torch.randn(max(400, n_users // 2), d, generator=gen)
    ie = torch.randn(n_items, d, generator=gen)
    ie[5] = ie[3]
    ie[n_items - 1] = ie[3]
    users = torch.randint(0, ue.shape[0], (n_users,), generator=gen)
    rowptr = flat = None
    if with_mask:
        lens = torch.randint(0, 40, (n_users,), generator=gen)
        lens[0] = 0
        cols = [torch.randperm(n_items, generator=gen)[:l].sort()[0] for l in lens.tolist()]
        rowptr = torch.tensor([0] + np.cumsum(lens.numpy()).tolist(), dtype=torch.int32).to(DEV)
        flat = torch.cat(cols).to(torch.int32).to(DEV)
    ue, ie, users = ue.to(DEV), ie.to(DEV), users.to(DEV)
    ids, vals = ops.score_mask_topk(ue, users, ie, k, rowptr, flat, n_splits=splits, return_scores=True)
    ids2, vals2 = ops.score_mask_topk(ue, users, ie, k, rowptr, flat, n_splits=splits, return_scores=True, simt=True)
    torch.cuda.synchronize()
    s = ue[users].double() @ ie.double().T
    if with_mask:
        mrows = torch.repeat_interleave(torch.arange(n_users, device=DEV), (rowptr[1:] - rowptr[:-1]).long())
        s[mrows, flat.long()] = -1e10
    want = torch.sort(s, dim=-1, descending=True, stable=True)[1][:, :k]
    same = (ids == want).float().mean().item()
    same_simt = (ids == ids2).float().mean().item()
    got_s = s.gather(1, ids)
    err = ((vals.double() - got_s).abs().max() / got_s.abs().clamp_max(1e9).max()).item()
    gap = (got_s - s.gather(1, want)).abs().max().item()
    v = vals
    eq = v[:, 1:] == v[:, :-1]
    ties_ok = bool((ids[:, 1:][eq] > ids[:, :-1][eq]).all())
    desc_ok = bool((v[:, 1:] <= v[:, :-1]).all())
    print(f"U={n_users} I={n_items} d={d} k={k} S={splits}: ids==f64 sort {same:.6f}  ==simt {same_simt:.6f} "
          f"score relerr {err:.2e}  max score gap at mismatches {gap:.2e}  ties_ok={ties_ok} desc_ok={desc_ok}",
          flush=True)
    return same > 0.999 and err < 1e-5 and ties_ok and desc_ok and gap < 1e-4


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ok = True
for args in [(64, 96, 64, 50, 1), (300, 1000, 64, 50, 4), (1000, 5000, 32, 50, None), (129, 777, 64, 20, 3),
             (9130, 7050, 64, 50, None), (128, 64, 64, 50, 1), (5, 70, 64, 50, 1)]:
    ok &= case(*args)
print("ALL OK" if ok else "MISMATCH", flush=True)

gen = torch.Generator().manual_seed(1)
for (nu, ni) in [(9130, 7050), (16384, 100000), (4096, 7050)]:
    ue = torch.randn(nu, 64, generator=gen).to(DEV)
    ie = torch.randn(ni, 64, generator=gen).to(DEV)
    users = torch.arange(nu, device=DEV)
    for S in (None, 1, 2, 4):
        t_tc = timeit(lambda: ops.score_mask_topk(ue, users, ie, 50, n_splits=S))
        print(f"time U={nu} I={ni} S={S}: tcgen05 {t_tc * 1e3:.1f} us  "
              f"({2 * nu * ni * 64 / t_tc / 1e9:.1f} TFLOP/s fp32-equivalent)", flush=True)
    t_simt = timeit(lambda: ops.score_mask_topk(ue, users, ie, 50, simt=True), iters=3)
    print(f"time U={nu} I={ni}: simt {t_simt * 1e3:.1f} us", flush=True)
