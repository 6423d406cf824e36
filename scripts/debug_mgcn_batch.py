"""Where does MGCN's batch-row step differ from the all-rows step? Per-parameter gradient differences and the
intermediate tables (all / side / content at the batch rows) for each dataset shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

DEV = "cuda:0"
for shape in sys.argv[1:] or ["tiny", "small", "sports"]:
    res = {}
    for mode in (False, True):
        env = bench.build_env(DEV, model_name="MGCN", shape=shape, overrides={"cuda_graph": False, "batch_rows": mode})
        m = env["model"]
        m.train()
        batch = bench.take_batches(env["train"], 1)[0]
        batch[0, 1::7] = batch[0, 0]
        batch[1, 2::5] = batch[1, 1]
        batch[2, ::3] = batch[1, ::3].roll(1)
        users, pos, neg = batch[0], batch[1], batch[2]
        with torch.no_grad():
            if mode:
                inter = m._forward_full(m.norm_adj, batch=(users, pos, neg))
            else:
                rows = torch.cat([users, m.n_users + pos, m.n_users + neg])
                inter = tuple(t[rows] for t in m._forward_full(m.norm_adj))
        m.zero_grad()
        loss = m.calculate_loss(batch)
        loss.backward()
        res[mode] = (float(loss.detach()), {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None},
                     inter)
    (l0, g0, i0), (l1, g1, i1) = res[False], res[True]
    print(f"[{shape}] B={int(users.numel())} loss all-rows {l0:.9g} batch-rows {l1:.9g} rel {abs(l0 - l1) / abs(l0):.2e}")
    for name, a, b in zip(("all_e", "side", "content"), i0, i1):
        print(f"  {name:8s} max|diff| {float((a - b).abs().max()):.3e} of max {float(a.abs().max()):.3e}")
    print("  keys only in one:", sorted(set(g0) ^ set(g1)))
    for k in sorted(set(g0) & set(g1)):
        scale = float(g0[k].abs().max().clamp_min(1e-12))
        diff = float((g0[k] - g1[k]).abs().max())
        print(f"  {k:40s} max|g| {scale:.3e} max|diff| {diff:.3e} rel {diff / scale:.2e}"
              + ("   <-- over 2e-5" if diff > 2e-5 * scale else ""))
