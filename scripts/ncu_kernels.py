"""One launch of every hot kernel at Baby-shaped sizes (after a warm-up launch): the command ncu wraps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ops, G, synth = bench.pkg("ops"), bench.pkg("graph"), bench.pkg("synth")
DEV = "cuda:0"
gen = torch.Generator().manual_seed(3)
I, U, d, F = 7050, 19445, 64, 4096
N = U + I
x = torch.randn(I, F, generator=gen).to(DEV)
W = (torch.randn(d, F, generator=gen) * 0.05).to(DEV)
b = torch.randn(d, generator=gen).to(DEV)
dy = torch.randn(I, d, generator=gen).to(DEV)
data = synth.make_dataset("baby", features=False)
u, i = data.split(0)
g = G.build_ui_graph(torch.from_numpy(u).to(DEV), torch.from_numpy(i).to(DEV), data.n_users, data.n_items, "f32")
X = torch.randn(N, d, device=DEV)
Y = torch.empty(N, d, device=DEV)
acc = torch.empty_like(X)
ue, ie = torch.randn(9130, d, device=DEV), torch.randn(I, d, device=DEV)
users = torch.arange(9130, device=DEV)
layers = [torch.nn.Linear(d, d, bias=bb).to(DEV) for bb in (True, False, True, False, True, True, True)]
ins = [torch.randn(N, d, device=DEV, requires_grad=True) for _ in range(4)]
masks = torch.nn.functional.dropout(torch.ones(3, N, d, device=DEV), 0.1)
drop_cnt = torch.tensor([3.0], dtype=torch.float64, device=DEV)
side, content = torch.randn(N, d, device=DEV, requires_grad=True), torch.randn(N, d, device=DEV, requires_grad=True)
bu = torch.randint(0, U, (2048,), device=DEV)
bi = torch.randint(0, I, (2048,), device=DEV)
img = torch.randn(I, d, device=DEV, requires_grad=True)
txt = torch.randn(I, d, device=DEV, requires_grad=True)
wsp = [torch.randn(d // 2 + 1, 2, device=DEV, requires_grad=True) for _ in range(3)]
gsp = [torch.randn(I, d, device=DEV) for _ in range(3)]
# round 2: low-rank table Adam (TMA-streamed), its norm pass, MGCN fuser, wide-d preference module, a graph
# whose embedding table (184 MB) does not fit L2
optim, par = bench.pkg("optim"), bench.pkg("parallel")
table = torch.nn.Parameter(x.clone())
opt = optim.FusedAdam([table], lr=1e-3)
g2 = torch.zeros(1, dtype=torch.float64, device=DEV)
mg = [torch.randn(N, d, device=DEV, requires_grad=True) for _ in range(8)]
w2 = torch.randn(1, d, device=DEV, requires_grad=True)
Nw, dw = 62420, 128
cw = [torch.randn(Nw, dw, device=DEV, requires_grad=True) for _ in range(9)]
xw = torch.randn(Nw, dw, device=DEV, requires_grad=True)
Ww, bw = (torch.randn(dw, dw, device=DEV) * 0.1).requires_grad_(True), torch.randn(dw, device=DEV, requires_grad=True)
su, si = synth.make_scaled_edges(DEV, 600_000, 120_000, 30_000_000)
big = G.build_ui_graph(su, si, 600_000, 120_000, "f64eps")
Xb = torch.randn(big.n_cols, d, device=DEV)
Yb = torch.empty(big.n_rows, d, device=DEV)
sbb = par.ShardedBipartite.from_local_edges(su, si, [0, 600_000], 0, 1, 600_000, 120_000, "f64eps", rt_block_users=196608)
yi = torch.empty(120_000, d, device=DEV)
for rep in range(2):
    table._mmrec_lowrank = ops.LowRankGrad(dy, W)
    optim.lowrank_sumsq(table, g2)
    opt.step()
    table._mmrec_lowrank = None
    a, s = ops.mgcn_fuse(mg[0], mg[1], w2, mg[2], mg[3], mg[4], mg[5], mg[6])
    (a.sum() + s.sum()).backward()
    a, s = ops.smore_combine(*cw)
    (a.sum() + s.sum()).backward()
    ops.dense_act(xw, Ww, bw, "sigmoid").sum().backward()
    ops.spmm_raw(big, Xb, Y=Yb)
    ops.spmm_blocked_raw(sbb.Rt_blocked, Xb[:600_000], yi)
    torch.autograd.backward(ops.spectrum_convolution(img, txt, *wsp, True), gsp)
    ops.gemm(x, True, W, True, I, d, F, b)
    ops.gemm(dy, False, x, False, d, F, I)
    ops.gemm(dy, True, W, False, I, F, d)
    ops.spmm_raw(g, X, Y=Y, acc_in=X, acc_out=acc)
    ops.score_mask_topk(ue, users, ie, 50)
    a, s = ops.smore_side(*ins, layers, masks)                       # explicit masks: the mma.sync forward
    (a.sum() + s.sum()).backward()
    a, s = ops.smore_side(*ins, layers, None, (0.5, 7, drop_cnt))     # in-kernel dropout: the tcgen05 forward (d = 64)
    (a.sum() + s.sum()).backward()
    l = ops.infonce_pair(side, content, U, bu, bi, 0.2)
    l.backward()
    torch.cuda.synchronize()
print("done")
