"""GPU parity tests added in round 2 (all through the C ABI):
* per-epoch edge dropout under CUDA-graph replay (stale-graph regression),
* low-rank feature-table gradients: fused tcgen05 Adam / sum-of-squares against the dense path,
  and the whole mirror-gradient trainer with and without them,
* MGCN's fused attention fuser (a10), the dense `full_sort_predict`, no silent library fallbacks.
"""
import numpy as np
import pytest
import torch

from conftest import pkg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


# ------------------------------------------------------------------ a4: dropout + graph replay
@pytest.mark.parametrize("model,over", [("FREEDOM", {"dropout": 0.8}), ("LayerGCN", {"dropout": 0.1})])
def test_edge_dropout_graph_replay_matches_eager(model, over):
    """layergcn.py:51-70 / freedom.py:130-145 replace `masked_adj` every epoch; a training step
    captured in a CUDA graph must follow it. Three epochs, graph replay vs eager, same RNG: the kept
    edge set changes per epoch, the graphs are re-captured per adjacency version, and the losses
    and final parameters agree."""
    from parity_util import make_env, golden_params
    import random
    out = {}
    for mode in (False, True):
        env = make_env(model, DEV, overrides=dict(over, cuda_graph=mode))
        m, train = env["model"], env["train"]
        m.load_state_dict({k: v.to(DEV) for k, v in golden_params(env["golden"]).items()})
        tr = pkg("trainer").Trainer(env["config"], m)
        assert tr.use_cuda_graph == mode
        torch.manual_seed(1234); torch.cuda.manual_seed(1234); random.seed(1234); np.random.seed(1234)
        losses, edge_sets, versions, graphed = [], [], [], 0
        for epoch in range(3):
            m.pre_epoch_processing()
            adj = m.masked_adj
            edge_sets.append(adj.col_idx.clone())
            versions.append(int(getattr(m, "graph_version", 0)))
            loss, _ = tr._train_epoch(train, epoch)
            tr.lr_scheduler.step()
            losses.append(loss)
            if mode:
                live = [k for k, e in tr._graphs.items() if e["graph"] is not None]
                graphed += len(live)
                assert all(k[2] == versions[-1] for k in tr._graphs), "a graph of an old adjacency survived"
        assert versions == sorted(set(versions)) and len(set(versions)) == 3        # bumped every epoch
        assert not torch.equal(edge_sets[0], edge_sets[1]) and not torch.equal(edge_sets[1], edge_sets[2])
        if mode:
            assert graphed > 0, "no step was ever replayed from a graph"
        out[mode] = (losses, {k: v.detach().clone() for k, v in m.state_dict().items()})
    np.testing.assert_allclose(out[True][0], out[False][0], rtol=1e-5)
    for k, v in out[False][1].items():
        assert rel(out[True][1][k], v) < 1e-4, k


# ------------------------------------------------------------------ low-rank table gradients
@pytest.mark.parametrize("rows,cols,d", [(7050, 4096, 64), (300, 128, 64), (1000, 384, 32), (2500, 256, 128),
                                         (129, 64, 64)])
def test_table_adam_lowrank_matches_dense_adam(rows, cols, d):
    """mmrec_table_adam_lowrank_f32 (gradient tiles rebuilt from dY, W on tcgen05, table and moments
    streamed by bulk copies) == materialise G = dY W with the library GEMM, then the multi-tensor
    Adam kernel. Same tensor-core products and the same float update: compared at 1e-6."""
    ops, optim = pkg("ops"), pkg("optim")
    gen = torch.Generator().manual_seed(rows + cols)
    X = torch.randn(rows, cols, generator=gen).to(DEV)
    W = (torch.randn(d, cols, generator=gen) * 0.05).to(DEV)
    a = torch.nn.Parameter(X.clone())
    b = torch.nn.Parameter(X.clone())
    oa, ob = optim.FusedAdam([a], lr=1e-2), optim.FusedAdam([b], lr=1e-2)
    # the dense gradient of the small shapes comes from the mma.sync GEMM (other summation order):
    # Adam's g / (|g| + eps) turns last-bit differences of near-zero gradients into 1e-5-class ones
    tol = 1e-6 if rows >= 1024 and cols >= 1024 else 3e-4
    for it in range(3):
        dY = (torch.randn(rows, d, generator=gen) * (10.0 ** (it - 1))).to(DEV)
        scale = 1.0 if it < 2 else -0.2
        a._mmrec_lowrank = ops.LowRankGrad(dY, W)
        b.grad = ops.gemm(dY, True, W, False, rows, cols, d)
        want_g2 = b.grad.double().pow(2).sum()
        g2 = torch.zeros(1, dtype=torch.float64, device=DEV)
        optim.lowrank_sumsq(a, g2)
        assert abs(float(g2) - float(want_g2)) / float(want_g2) < 1e-5
        oa.step(grad_scale=scale); ob.step(grad_scale=scale)
        oa.zero_grad(); ob.zero_grad()
        assert a._mmrec_lowrank is None
        assert rel(a.detach(), b.detach()) < tol
        assert rel(oa.state[a]["exp_avg"], ob.state[b]["exp_avg"]) < tol
        assert rel(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"]) < tol
        want_p2 = a.detach().double().pow(2).sum()
        assert abs(float(oa.state[a]["sumsq"]) - float(want_p2)) / float(want_p2) < 1e-5
    # the update count lives once per group and is ticked once per step
    assert float(oa.param_groups[0]["hyper"][1]) == 3.0


def test_table_project_lowrank_forward_backward_and_mirror_point():
    """ops.table_project == F.linear on the table (forward, dW, db), leaves the factors of the table
    gradient on the parameter, and evaluates the mirror point X - c dY1 W1 without writing it."""
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(5)
    I, F, d = 1500, 384, 64
    emb = torch.nn.Embedding.from_pretrained(torch.randn(I, F, generator=gen).to(DEV), freeze=False)
    W = torch.nn.Parameter((torch.randn(d, F, generator=gen) * 0.05).to(DEV))
    b = torch.nn.Parameter(torch.randn(d, generator=gen).to(DEV))
    Gy = torch.randn(I, d, generator=gen).to(DEV)
    y = ops.table_project(emb, W, b)
    (y * Gy).sum().backward()
    Xr, Wr, br = (t.detach().double().requires_grad_(True) for t in (emb.weight, W, b))
    yr = torch.nn.functional.linear(Xr, Wr, br)
    (yr * Gy.double()).sum().backward()
    assert rel(y, yr) < 1e-6 and rel(W.grad, Wr.grad) < 1e-6 and rel(b.grad, br.grad) < 1e-6
    assert emb.weight.grad is None
    lr = emb.weight._mmrec_lowrank
    assert rel(lr.dense(), Xr.grad) < 1e-6
    # mirror point
    coef = torch.tensor([0.37], device=DEV)
    dY1, W1 = lr.dY, lr.W.clone()
    emb.weight._mmrec_lowrank = None
    emb.weight._mmrec_delta = (coef, dY1, W1)
    W.grad = b.grad = None
    y2 = ops.table_project(emb, W, b)
    (y2 * Gy).sum().backward()
    X2 = (emb.weight.detach().double() - 0.37 * dY1.double() @ W1.double())
    W2 = W.detach().double().requires_grad_(True)
    y2r = torch.nn.functional.linear(X2, W2, b.detach().double())
    (y2r * Gy.double()).sum().backward()
    assert rel(y2, y2r) < 1e-6 and rel(W.grad, W2.grad) < 1e-6
    emb.weight._mmrec_delta = None
    with pytest.raises(RuntimeError):                       # a second backward would have to accumulate
        (ops.table_project(emb, W, b) * Gy).sum().backward()


@pytest.mark.parametrize("graph", [False, True])
def test_smore_trainer_lowrank_tables_match_dense_tables(graph):
    """The mirror-gradient trainer (trainer.py:268-335) with the table gradients kept as factors
    (virtual mirror point, fused table Adam, Gram-free norms) follows the dense implementation."""
    from parity_util import make_env, golden_params
    out = {}
    for lowrank in (False, True):
        env = make_env("SMORE", DEV, overrides={"cuda_graph": graph, "dropout_rate": 0.0,
                                                 "lowrank_table_grad": lowrank})
        m, train = env["model"], env["train"]
        m.load_state_dict({k: v.to(DEV) for k, v in golden_params(env["golden"]).items()})
        tr = pkg("trainer").Trainer(env["config"], m)
        assert m.lowrank_table_grad == lowrank
        losses = []
        for epoch in range(3):
            m.pre_epoch_processing()
            loss, _ = tr._train_epoch(train, epoch)
            tr.lr_scheduler.step()
            losses.append(loss)
        if lowrank:
            assert "sumsq" in tr.optimizer.state[m.image_embedding.weight], "the low-rank path never ran"
        out[lowrank] = (losses, {k: v.detach().clone() for k, v in m.state_dict().items()}, float(m._alpha_eff))
    np.testing.assert_allclose(out[True][0], out[False][0], rtol=2e-6)
    assert abs(out[True][2] - out[False][2]) <= 1e-5 * abs(out[False][2])
    for k, v in out[False][1].items():
        assert rel(out[True][1][k], v) < 2e-5, k


def test_fused_adam_state_dict_round_trip_with_torch_adam():
    """state_dict carries torch's per-parameter `step`; loading a torch.optim.Adam state resumes the
    bias correction where it stopped (it silently restarted from zero)."""
    optim = pkg("optim")
    gen = torch.Generator().manual_seed(3)
    shapes = [(257, 64), (64,)]
    a = [torch.nn.Parameter(torch.randn(*s, generator=gen).to(DEV)) for s in shapes]
    b = [torch.nn.Parameter(t.detach().clone()) for t in a]
    ta = torch.optim.Adam(a, lr=1e-2)
    tb = torch.optim.Adam(b, lr=1e-2)
    grads = [[torch.randn(*s, generator=gen).to(DEV) for s in shapes] for _ in range(5)]
    for it in range(3):
        for x, y, g in zip(a, b, grads[it]):
            x.grad, y.grad = g.clone(), g.clone()
        ta.step(); tb.step()
    fa = optim.FusedAdam(a, lr=1e-2)
    fa.load_state_dict(ta.state_dict())
    assert float(fa.param_groups[0]["hyper"][1]) == 3.0
    for it in range(3, 5):
        for x, y, g in zip(a, b, grads[it]):
            x.grad, y.grad = g.clone(), g.clone()
        fa.step(); tb.step()
    for x, y in zip(a, b):
        assert rel(x.detach(), y.detach()) < 1e-6
    sd = fa.state_dict()
    assert all(float(st["step"]) == 5.0 for st in sd["state"].values())


# ------------------------------------------------------------------ a10: MGCN fuser
@pytest.mark.parametrize("n,d", [(53955, 64), (1, 64), (777, 32), (300, 128)])
def test_mgcn_fuse_forward_backward(n, d):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(n + d)
    names = ("Hi", "Ht", "Ei", "Et", "Pi", "Pt", "C")
    t = {k: torch.randn(n, d, generator=gen) for k in names}
    t["Hi"], t["Ht"] = torch.tanh(t["Hi"]), torch.tanh(t["Ht"])
    t["Pi"], t["Pt"] = torch.sigmoid(t["Pi"]), torch.sigmoid(t["Pt"])
    w2 = torch.randn(1, d, generator=gen) * 0.3
    ga, gs = torch.randn(n, d, generator=gen), torch.randn(n, d, generator=gen)

    def ref(dtype):
        x = {k: v.to(dtype).requires_grad_(True) for k, v in t.items()}
        w = w2.to(dtype).requires_grad_(True)
        att = torch.softmax(torch.cat([x["Hi"] @ w.t(), x["Ht"] @ w.t()], dim=-1), dim=-1)      # mgcn.py:188-190
        common = att[:, 0:1] * x["Ei"] + att[:, 1:2] * x["Et"]
        side = (x["Pi"] * (x["Ei"] - common) + x["Pt"] * (x["Et"] - common) + common) / 3       # mgcn.py:192-203
        all_e = x["C"] + side
        ((all_e * ga.to(dtype)).sum() + (side * gs.to(dtype)).sum()).backward()
        return all_e, side, x, w

    want_all, want_side, xr, wr = ref(torch.float64)
    x = {k: v.to(DEV).requires_grad_(True) for k, v in t.items()}
    w = w2.to(DEV).requires_grad_(True)
    all_e, side = ops.mgcn_fuse(x["Hi"], x["Ht"], w, x["Ei"], x["Et"], x["Pi"], x["Pt"], x["C"])
    ((all_e * ga.to(DEV)).sum() + (side * gs.to(DEV)).sum()).backward()
    assert rel(all_e, want_all) < 1e-6 and rel(side, want_side) < 1e-6
    for k in names:
        assert rel(x[k].grad, xr[k].grad) < 5e-6, k
    assert rel(w.grad, wr.grad) < 5e-6
    # only one of the two outputs used downstream (g_side = None)
    x2 = {k: v.to(DEV).requires_grad_(True) for k, v in t.items()}
    a2, _ = ops.mgcn_fuse(x2["Hi"], x2["Ht"], w2.to(DEV), x2["Ei"], x2["Et"], x2["Pi"], x2["Pt"], x2["C"])
    (a2 * ga.to(DEV)).sum().backward()
    assert torch.equal(x2["C"].grad, ga.to(DEV))


# ------------------------------------------------------------------ no silent library paths
def test_linear_has_no_cublas_fallback():
    ops = pkg("ops")
    x = torch.randn(40, 64, device=DEV)
    with pytest.raises(RuntimeError):
        ops.linear(x, torch.randn(1, 64, device=DEV))            # Linear(d, 1): not a GEMM shape of this library
    with pytest.raises(RuntimeError):
        ops.linear(x[:, :30], torch.randn(8, 30, device=DEV))


@pytest.mark.parametrize("n_items", [7050, 96, 1001])
def test_full_sort_predict_dense_scores_on_library_gemm(n_items):
    """layergcn.py:186-188 / smore.py:419-422: the API-compatible [Bu, n_items] score matrix."""
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(n_items)
    u = torch.randn(333, 64, generator=gen).to(DEV)
    v = torch.randn(n_items, 64, generator=gen).to(DEV)
    s = ops.score_matrix(u, v)
    assert tuple(s.shape) == (333, n_items)
    assert rel(s, u.double() @ v.double().t()) < 1e-6


def test_model_full_sort_predict_matches_fused_topk():
    from parity_util import make_env, golden_params
    env = make_env("LayerGCN", DEV)
    m = env["model"]
    m.load_state_dict({k: v.to(DEV) for k, v in golden_params(env["golden"]).items()})
    m.eval()
    users = torch.arange(0, 64, device=DEV)
    scores = m.full_sort_predict((users,))
    ids = m.full_sort_topk(users, 10)
    want = torch.sort(scores, dim=1, descending=True, stable=True)[1][:, :10]
    assert torch.equal(ids, want)


# ------------------------------------------------------------------ column-blocked SpMM (X > L2)
def test_column_blocked_spmm_and_shard_local_bipartite_build():
    """graph.ColumnBlockedCSR (Y = A_0 X_0, Y += A_b X_b) == the unblocked SpMM, and a
    ShardedBipartite built from a rank's own edges (with and without user blocks) propagates like
    the symmetric full graph."""
    G, ops, par, synth = pkg("graph"), pkg("ops"), pkg("parallel"), pkg("synth")
    U, I, E, d = 20000, 3000, 400000, 64
    su, si = synth.make_scaled_edges(DEV, U, I, E, block=4096)
    full = G.build_ui_graph(su, si, U, I, "f64eps")
    gen = torch.Generator().manual_seed(9)
    X = torch.randn(U + I, d, generator=gen).to(DEV)
    want = ops.propagate_mean(full, X, 3)
    bounds = np.array([0, U], dtype=np.int64)
    for blk in (None, 3000, 7001):
        sb = par.ShardedBipartite.from_local_edges(su, si, bounds, 0, 1, U, I, "f64eps", rt_block_users=blk)
        assert (sb.Rt_blocked is not None) == (blk is not None)
        if blk is not None:
            assert len(sb.Rt_blocked.blocks) == -(-U // blk) and sb.Rt_blocked.nnz == su.numel()
            # the blocked operator alone against the plain one
            Rt = G.csr_from_coo(si, su, sb.R.vals.new_ones(su.numel()), I, U, with_transpose=False)
            ones = G.ColumnBlockedCSR.from_col_sorted_coo(si, su, sb.R.vals.new_ones(su.numel()), I, U, blk)
            y0 = torch.empty(I, d, device=DEV)
            y1 = torch.full((I, d), float("nan"), device=DEV)
            ops.spmm_raw(Rt, X[:U].contiguous(), Y=y0)
            ops.spmm_blocked_raw(ones, X[:U].contiguous(), y1)
            assert rel(y1, y0) < 1e-6
        ou, oi = par.bipartite_propagate_mean(sb, X[:U].contiguous(), X[U:].contiguous(), 3)
        assert rel(ou, want[:U]) < 2e-6 and rel(oi, want[U:]) < 2e-6
    # values of the shard-local build are the full graph's, bit for bit
    r, c, v = full.to_torch_coo()
    sb = par.ShardedBipartite.from_local_edges(su, si, bounds, 0, 1, U, I, "f64eps")
    r2, c2, v2 = sb.R.to_torch_coo()
    m = r < U
    assert np.array_equal(r[m], r2) and np.array_equal(c[m] - U, c2) and np.array_equal(v[m].view(np.uint32), v2.view(np.uint32))


# ------------------------------------------------------------------ f3: device negative sampler
def test_device_negative_sampler_matches_numpy_restatement_and_rule():
    """mmrec_neg_sample_counter == oracle/sampler.py bit for bit; no negative is in its user's history;
    the draws are uniform over the items outside it; batches are reproducible from (seed, step)."""
    from oracle import sampler as osampler
    data_m, synth = pkg("data"), pkg("synth")
    U, I = 3000, 500
    su, si = synth.make_scaled_edges(DEV, U, I, 60000, block=1024)
    dl = data_m.DeviceTrainLoader(su, si, U, I, batch_size=4096, seed=7)
    rp, hc = dl.hist_rowptr.cpu().numpy(), dl.hist_cols.cpu().numpy()
    seen = []
    for step, batch in enumerate(dl):
        assert batch.shape[0] == 3 and batch.dtype == torch.int64 and batch.is_cuda
        u, pos, neg = (t.cpu().numpy() for t in batch)
        want = osampler.neg_sample_counter(u, None, I, rp, hc, 7, step)
        assert np.array_equal(neg, want)
        key = np.sort(su.cpu().numpy() * I + si.cpu().numpy())
        assert not np.isin(u * I + neg, key).any() and (neg >= 0).all()
        assert np.isin(u * I + pos, key).all()
        seen.append(neg)
        if step == 3:
            break
    again = dl.sample_negatives(batch[0], 3)
    assert torch.equal(again, batch[2])                      # stateless: (seed, step, position) -> same draw
    counts = np.bincount(np.concatenate(seen), minlength=I)
    assert counts.min() > 0 and counts.max() < 4 * counts.mean()      # uniform over items, not popularity-biased
    # a user whose history is (almost) everything: the rejection loop still terminates
    hist_u = torch.zeros(I - 1, dtype=torch.int64, device=DEV)
    hist_i = torch.arange(I - 1, device=DEV)
    dl2 = data_m.DeviceTrainLoader(hist_u, hist_i, 1, I, batch_size=64, seed=1, max_draws=4096)
    neg2 = dl2.sample_negatives(torch.zeros(64, dtype=torch.int64, device=DEV), 0)
    assert bool(((neg2 == I - 1) | (neg2 == -1)).all()) and int((neg2 == I - 1).sum()) > 0


# ------------------------------------------------------------------ a10 at d = 128: row part of the preference module
@pytest.mark.parametrize("n,d,drop", [(62420, 128, 0.1), (1, 128, 0.0), (777, 64, 0.2), (300, 32, 0.0)])
def test_smore_combine_forward_backward(n, d, drop):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(n + d)
    names = ("zv", "zt", "V", "T", "F", "C", "gi", "gt", "gf")
    t = {k: torch.randn(n, d, generator=gen) for k in names}
    for k in ("gi", "gt", "gf"):
        t[k] = torch.sigmoid(t[k])
    masks = None
    if drop > 0:
        masks = (torch.rand(3, n, d, generator=gen) >= drop).float() / (1 - drop)
    ga, gs = torch.randn(n, d, generator=gen), torch.randn(n, d, generator=gen)
    x = {k: v.double().requires_grad_(True) for k, v in t.items()}
    m = masks.double() if masks is not None else torch.ones(3, n, d, dtype=torch.float64)
    agg_i = torch.softmax(x["zv"], dim=-1) * x["V"]                                  # smore.py:324-325
    agg_t = torch.softmax(x["zt"], dim=-1) * x["T"]
    side_r = torch.mean(torch.stack([x["gi"] * m[0] * agg_i, x["gt"] * m[1] * agg_t, x["gf"] * m[2] * x["F"]]), dim=0)
    all_r = x["C"] + side_r                                                            # smore.py:335-341
    ((all_r * ga.double()).sum() + (side_r * gs.double()).sum()).backward()
    y = {k: v.to(DEV).requires_grad_(True) for k, v in t.items()}
    all_e, side = ops.smore_combine(y["zv"], y["zt"], y["V"], y["T"], y["F"], y["C"], y["gi"], y["gt"], y["gf"],
                                    None if masks is None else masks.to(DEV))
    ((all_e * ga.to(DEV)).sum() + (side * gs.to(DEV)).sum()).backward()
    assert rel(all_e, all_r) < 1e-6 and rel(side, side_r) < 1e-6
    for k in names:
        assert rel(y[k].grad, x[k].grad) < 5e-6, k


def test_id_range_checks_catch_out_of_bounds_indices(monkeypatch):
    """MMREC_CHECK_IDS: the gather / scatter entry points validate their ids (the stand-in for
    compute-sanitizer's memcheck on the id arguments)."""
    ops = pkg("ops")
    monkeypatch.setattr(ops, "CHECK_IDS", True)
    emb = torch.randn(50, 64, device=DEV, requires_grad=True)
    ok = torch.tensor([0, 9], device=DEV)
    ops.bpr_table(emb, 10, ok, ok, ok)
    with pytest.raises(IndexError):
        ops.bpr_table(emb, 10, torch.tensor([0, 10], device=DEV), ok, ok)          # user id == n_users
    with pytest.raises(IndexError):
        ops.bpr_table(emb, 10, ok, torch.tensor([0, 40], device=DEV), ok)          # item id == n_items
    with pytest.raises(IndexError):
        ops.score_mask_topk(emb[:10].detach(), torch.tensor([-1], device=DEV), emb[10:].detach(), 5)
