"""One launch of the batch-row kernels at Baby-shaped sizes (after a warm-up launch): the command ncu wraps.
gather / scatter of the batch rows of four node tables, the batch views (user rows R x' for the batch users only)
and the preference module on the 3 B = 6 144 compact rows (forward on tcgen05 + mma.sync backward)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ops = bench.pkg("ops")
DEV = "cuda:0"
env = bench.build_env(DEV, overrides={"cuda_graph": False})
m = env["model"]
batch = bench.take_batches(env["train"], 1)[0]
users, pos, neg = batch[0], batch[1], batch[2]
U, I, d = m.n_users, m.n_items, m.embedding_dim
N = U + I
tables = [torch.randn(N, d, device=DEV, requires_grad=True) for _ in range(4)]
items = [torch.randn(I, d, device=DEV, requires_grad=True) for _ in range(3)]
content = torch.randn(N, d, device=DEV, requires_grad=True)
layers = [torch.nn.Linear(d, d, bias=bb).to(DEV) for bb in (True, False, True, False, True, True, True)]
cnt = torch.tensor([3.0], dtype=torch.float64, device=DEV)
for rep in range(2):
    ids, out = ops.gather_batch_rows(tables, users, pos, neg, U)
    a, s = ops.smore_side(*out, layers, None, (0.1, 7, cnt, ids, N))
    (a.sum() + s.sum()).backward()
    ids, c, views = ops.gather_batch_views(m.R, items, content, users, pos, neg)
    (c.sum() + sum(v.sum() for v in views)).backward()
    torch.cuda.synchronize()
print("done")
