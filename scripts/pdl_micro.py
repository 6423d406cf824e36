"""Programmatic dependent launch on the SpMM chain: a 4-layer propagate_mean (forward) and its Horner backward on
the Baby-shaped UI graph, captured in a CUDA graph and replayed; MMREC_PDL=0 against 1 in one process (the switch
is read at launch = capture time). Also checks that both settings give bit-identical results."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "recommendar-systems_b200"
synth, G, ops = (importlib.import_module(f"{PKG}.{m}") for m in ("synth", "graph", "ops"))
dev = "cuda:0"
shape = sys.argv[1] if len(sys.argv) > 1 else "baby"
d = synth.make_dataset(shape, features=False)
u, i = d.split(0)
g = G.build_ui_graph(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev), d.n_users, d.n_items, "f32")
X = torch.randn(g.n_cols, 64, device=dev)
outs = {}
for pdl in ("0", "1", "0", "1"):
    os.environ["MMREC_PDL"] = pdl
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            y = ops._PropagateMean.apply(X, g, 4); z = ops._horner(g, y, 4, 0.2)
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            y = ops._PropagateMean.apply(X, g, 4); z = ops._horner(g, y, 4, 0.2)
        for _ in range(5): gr.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100): gr.replay()
        b.record(); s.synchronize()
    us = a.elapsed_time(b) * 10 / 8
    outs[pdl] = (y.clone(), z.clone())
    print(f"{shape}: MMREC_PDL={pdl}: {us:.2f} us per SpMM launch (8 chained launches per replay, graph replay)", flush=True)
print("bit-identical:", torch.equal(outs["0"][0], outs["1"][0]) and torch.equal(outs["0"][1], outs["1"][1]))
