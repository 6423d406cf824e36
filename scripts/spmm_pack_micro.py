"""SpMM task packing A/B (graph.build_tasks(pack=...)): Baby / Sports UI graphs, a kNN-like item graph, R and R^T
views, the three-problem launch; the >L2 600k x 120k graph. Checks bit-equality of packed vs one-row task lists, times
launches inside a CUDA graph of 40 back-to-back launches (the in-step regime: L2 warm, no launch gaps)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "recommendar-systems_b200"
synth, G, ops = (importlib.import_module(f"{PKG}.{m}") for m in ("synth", "graph", "ops"))
dev = "cuda:0"
D = int(os.environ.get("D", "64"))


def clone_graph(g, pack):
    h = G.CSRGraph(g.row_ptr, g.col_idx, g.vals, g.n_rows, g.n_cols, col_offset=g.col_offset, pack=pack)
    return h


def timed(fn, reps=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        a.record()
        gr.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best * 1e3


def case(name, g, d=D, fused=True):
    X = torch.randn(g.n_cols, d, device=dev)
    res = {}
    outs = {}
    for pack in (False, True):
        h = clone_graph(g, pack)
        Y = torch.empty(g.n_rows, d, device=dev)
        acc_in = torch.randn(g.n_rows, d, device=dev)
        acc = torch.empty(g.n_rows, d, device=dev)
        if fused:
            fn = lambda: ops.spmm_raw(h, X, Y=Y, acc_in=acc_in, acc_out=acc)
        else:
            fn = lambda: ops.spmm_raw(h, X, Y=Y)
        res[pack] = (timed(fn), h.n_tasks)
        outs[pack] = (Y.clone(), acc.clone() if fused else None)
    same = torch.equal(outs[False][0], outs[True][0]) and (not fused or torch.equal(outs[False][1], outs[True][1]))
    byt = g.algorithmic_bytes(d) + (4 * d * g.n_rows // 4 if fused else 0)
    print(f"{name:34s} rows {g.n_rows:7d} nnz {g.nnz:9d} | one-row {res[False][0]:7.2f} us ({res[False][1]} tasks) | "
          f"packed(W={G.PACK_WINDOW}) {res[True][0]:7.2f} us ({res[True][1]} tasks) | x{res[False][0] / res[True][0]:.2f} | "
          f"bit-identical {same} | {byt / res[True][0] / 1e3:.0f} GB/s algorithmic", flush=True)


def ui(shape):
    d = synth.make_dataset(shape, features=False)
    u, i = d.split(0)
    g = G.build_ui_graph(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev), d.n_users, d.n_items, "f32")
    return d, g


which = sys.argv[1:] or ["baby", "sports", "knn", "big"]
if "baby" in which:
    d, g = ui("baby")
    case("baby UI (fused layer-sum)", g)
    case("baby UI (plain)", g, fused=False)
    R, Rt = G.ui_blocks(g)
    case("baby R (users <- items)", R, fused=False)
    case("baby R^T (items <- users)", Rt, fused=False)
    case("baby UI d=128", g, d=128)
if "sports" in which:
    d, g = ui("sports")
    case("sports UI (fused layer-sum)", g)
if "knn" in which:
    I, k = 7050, 10
    rows = torch.arange(I, device=dev).repeat_interleave(k)
    cols = torch.randint(0, I, (I * k,), device=dev)
    vals = torch.rand(I * k, device=dev)
    g = G.csr_from_coo(rows, cols, vals, I, I, with_transpose=False)
    case("kNN item graph 7050 x 10", g, fused=False)
if "big" in which:
    su, si = synth.make_scaled_edges(dev, 600_000, 120_000, 30_000_000)
    g = G.build_ui_graph(su, si, 600_000, 120_000, "f64eps")
    case("600k x 120k (> L2)", g)
