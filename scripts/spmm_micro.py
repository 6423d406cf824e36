"""SpMM micro-benchmark: Baby-shaped UI graph (L2-resident) and a >L2 scaled graph."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "recommendar-systems_b200"
synth, G, ops = (importlib.import_module(f"{PKG}.{m}") for m in ("synth", "graph", "ops"))
which = sys.argv[1] if len(sys.argv) > 1 else "baby"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda:0"
if which == "baby":
    d = synth.make_dataset("baby", features=False)
    u, i = d.split(0)
    g = G.build_ui_graph(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev), d.n_users, d.n_items, "f32")
else:
    su, si = synth.make_scaled_edges(dev, 600_000, 120_000, 30_000_000)
    g = G.build_ui_graph(su, si, 600_000, 120_000, "f64eps")
X = torch.randn(g.n_cols, 64, device=dev)
Y = torch.empty(g.n_rows, 64, device=dev)
acc = torch.empty_like(X)
for _ in range(3):
    ops.spmm_raw(g, X, Y=Y, acc_in=X, acc_out=acc)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    ops.spmm_raw(g, X, Y=Y, acc_in=X, acc_out=acc)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
byt = g.algorithmic_bytes(64) + 8 * 64 * g.n_rows
print(f"{which}: rows {g.n_rows} nnz {g.nnz} tasks {g.n_tasks} heavy parts {g.total_parts} | "
      f"{ms * 1e3:.1f} us/launch back-to-back (warm L2), {byt / ms / 1e6:.0f} GB/s algorithmic")
