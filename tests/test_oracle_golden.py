"""Pin the oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden, pkg
from oracle import build, graph, models, ops

CFG = {
    "LightGCN": {"n_layers": 4, "reg_weight": 1e-2},
    "LayerGCN": {"n_layers": 4, "reg_weight": 1e-2},
    "FREEDOM": {"n_mm_layers": 1, "n_ui_layers": 2, "knn_k": 10, "mm_image_weight": 0.1,
                "reg_weight": 1e-3},
    "MGCN": {"n_ui_layers": 2, "n_layers": 1, "knn_k": 10, "reg_weight": 1e-4, "cl_loss": 0.01,
             "train_batch_size": 512},
    "SMORE": {"n_ui_layers": 4, "n_layers": 1, "image_knn_k": 20, "text_knn_k": 15,
              "reg_weight": 1e-5, "cl_loss": 0.01, "train_batch_size": 512},
}
TAGS = {"LightGCN": "tiny_lightgcn", "LayerGCN": "tiny_layergcn", "FREEDOM": "tiny_freedom",
        "MGCN": "tiny_mgcn", "SMORE": "tiny_smore"}
ADJ_KEY = {"LightGCN": "norm_adj_matrix", "LayerGCN": "norm_adj_matrix", "FREEDOM": "norm_adj",
           "MGCN": "norm_adj", "SMORE": "norm_adj"}


def sort_coo(r, c, v):
    o = np.lexsort((c, r))
    return r[o], c[o], v[o]


# --------------------------------------------------------------------------- micro vectors
def test_micro_adjacency_recipes():
    g = golden("micro")
    u = np.array([0, 0, 1, 2]); i = np.array([0, 1, 1, 0])
    r, c, v = graph.norm_adj_f64eps(u, i, 3, 2)
    assert np.array_equal(np.vstack([r, c]), g["layergcn_adj_idx"])
    assert np.array_equal(v.view(np.uint32), g["layergcn_adj_val"].view(np.uint32))
    assert float.hex(float(v[2])) == "0x1.6a09e40000000p-1"
    r, c, v = graph.norm_adj_f32(u, i, 3, 2)
    assert np.array_equal(np.vstack([r, c]), g["mgcn_adj_idx"])
    assert np.array_equal(v.view(np.uint32), g["mgcn_adj_val"].view(np.uint32))
    assert float.hex(float(v[2])) == "0x1.6a09e60000000p-1"
    rr, rc, rv = graph.r_block(r, c, v, 3)
    assert np.array_equal(np.vstack([rr, rc]), g["mgcn_R_idx"])
    assert np.array_equal(rv.view(np.uint32), g["mgcn_R_val"].view(np.uint32))
    ev = graph.edge_norm_f32(u, i, 3, 2)
    assert np.array_equal(ev.view(np.uint32), g["edge_norm_val"].view(np.uint32))
    feats = torch.tensor([[1., 0, 0], [.9, .1, 0], [0, 1., 0]])
    r, c, v = graph.freedom_knn_adj(feats, 2)
    assert np.array_equal(np.vstack([r, c]), g["freedom_knn_idx"])
    assert np.array_equal(v.view(np.uint32), g["freedom_knn_val"].view(np.uint32))


def test_micro_forward_loss_metrics_spectrum():
    g = golden("micro")
    A = graph.to_torch_csr(g["layergcn_adj_idx"][0], g["layergcn_adj_idx"][1],
                           g["layergcn_adj_val"], (5, 5))
    x0 = torch.tensor([[1., 0], [0, 1], [1, 1], [1, 2], [2, 1]])
    out = ops.layergcn_propagate(A, x0, 2)
    np.testing.assert_allclose(out[:3].numpy(), g["layergcn_fwd_user"], rtol=1e-6)
    np.testing.assert_allclose(out[3:].numpy(), g["layergcn_fwd_item"], rtol=1e-6)
    np.testing.assert_allclose(ops.propagate_mean(A, x0, 2).numpy(), g["lightgcn_mean"], rtol=1e-6)
    users, pos, neg = torch.tensor([0, 2]), torch.tensor([1, 0]), torch.tensor([0, 1])
    loss = ops.bpr_sum(out[:3][users], out[3:][pos], out[3:][neg])
    np.testing.assert_allclose(loss.item(), g["layergcn_bpr_sum"], rtol=1e-6)
    hits = np.array([[1, 0, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 0]], dtype=bool)
    pos_len = np.array([2, 1, 7])
    for name in ("recall", "ndcg", "precision", "map"):
        np.testing.assert_allclose(ops.METRICS[name](hits, pos_len), g["metric_" + name], rtol=1e-12)
    w = torch.from_numpy(g["spec_w"])
    ic, tc, fc = ops.spectrum_convolution(torch.from_numpy(g["spec_x"]), torch.from_numpy(g["spec_y"]),
                                          w, w, w, True)
    np.testing.assert_allclose(ic.numpy(), g["spec_image_conv"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(tc.numpy(), g["spec_text_conv"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(fc.numpy(), g["spec_fusion_conv"], rtol=1e-6, atol=1e-6)
    assert float(g["cos_tiny"][0]) == pytest.approx(0.01)             # torch 2.x cosine semantics


# --------------------------------------------------------------------------- tiny dataset
def _graphs(model, tiny_data, tiny_train, keep_idx=None, dtype=torch.float32):
    u, i = tiny_train
    img = tiny_data.image_feat if model in ("FREEDOM", "MGCN", "SMORE") else None
    return build.build_graphs(model, u, i, tiny_data.n_users, tiny_data.n_items, img,
                              tiny_data.text_feat, CFG[model], dtype=dtype, keep_idx=keep_idx)


@pytest.mark.parametrize("model", list(TAGS))
def test_adjacency_bit_exact(model, tiny_data, tiny_train):
    g = golden(TAGS[model])
    G, parts = _graphs(model, tiny_data, tiny_train)
    idx, val = g[f"adj/{ADJ_KEY[model]}/idx"], g[f"adj/{ADJ_KEY[model]}/val"]
    r, c, v = parts["norm_adj"]
    assert np.array_equal(np.vstack([r, c]), idx)                     # same order, too
    assert np.array_equal(v.view(np.uint32), val.view(np.uint32))
    if "R" in parts:
        r, c, v = parts["R"]
        assert np.array_equal(np.vstack([r, c]), g["adj/R/idx"])
        assert np.array_equal(v.view(np.uint32), g["adj/R/val"].view(np.uint32))
    names = {"FREEDOM": [("mm_adj", "mm_adj")],
             "MGCN": [("image_adj", "image_original_adj"), ("text_adj", "text_original_adj")],
             "SMORE": [("image_adj", "image_original_adj"), ("text_adj", "text_original_adj"),
                       ("fusion_adj", "fusion_adj")]}.get(model, [])
    for ours, theirs in names:
        r, c, v = sort_coo(*parts[ours])
        gi, gv = g[f"adj/{theirs}/idx"], g[f"adj/{theirs}/val"]
        t = torch.sparse_coo_tensor(gi, torch.from_numpy(gv)).coalesce()
        assert np.array_equal(np.vstack([r, c]), t.indices().numpy())
        assert np.array_equal(v.view(np.uint32), t.values().numpy().view(np.uint32))


@pytest.mark.parametrize("tag", ["tiny_layergcn_drop", "tiny_freedom"])
def test_edge_dropout_renormalisation(tag, tiny_data, tiny_train):
    """a4: given the reference's kept edges, the re-normalised mirrored adjacency is bit-exact."""
    g = golden(tag)
    u, i = graph.edge_list(*tiny_train)
    assert np.array_equal(g["edge_indices"], np.vstack([u, i]))
    ev = graph.edge_norm_f32(u, i, tiny_data.n_users, tiny_data.n_items)
    assert np.array_equal(ev.view(np.uint32), g["edge_values"].view(np.uint32))
    midx, mval = g["adj/masked_adj/idx"], g["adj/masked_adj/val"]
    keep = midx.shape[1] // 2
    # recover keep_idx from the kept (u, i) pairs
    key = u * tiny_data.n_items + i
    kept_key = midx[0, :keep] * tiny_data.n_items + (midx[1, :keep] - tiny_data.n_users)
    order = np.argsort(key)
    keep_idx = order[np.searchsorted(key[order], kept_key)]
    r, c, v = graph.masked_adj(u, i, keep_idx, tiny_data.n_users, tiny_data.n_items)
    assert np.array_equal(np.vstack([r, c]), midx)
    assert np.array_equal(v.view(np.uint32), mval.view(np.uint32))


def _params(g, dtype=torch.float32, grad=False):
    P = {}
    for k in g.files:
        if k.startswith("param0/"):
            t = torch.from_numpy(g[k]).to(dtype)
            P[k[len("param0/"):]] = t.requires_grad_(grad)
    return P


@pytest.mark.parametrize("model", list(TAGS))
def test_forward_loss_grads_match_reference(model, tiny_data, tiny_train):
    g = golden(TAGS[model])
    G, _ = _graphs(model, tiny_data, tiny_train)
    if model in ("LayerGCN", "FREEDOM"):
        midx, mval = g["adj/masked_adj/idx"], g["adj/masked_adj/val"]
        n = tiny_data.n_users + tiny_data.n_items
        G["masked_adj"] = graph.to_torch_csr(midx[0], midx[1], mval, (n, n))
    P = _params(g, grad=True)
    cfg = CFG[model]
    with torch.no_grad():
        ue, ie = models.FORWARD[model](P, G, cfg)
    np.testing.assert_allclose(ue.numpy(), g["eval_user_emb"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(ie.numpy(), g["eval_item_emb"], rtol=2e-5, atol=1e-7)
    batch = torch.from_numpy(g["batch0"])
    loss = models.LOSS[model](P, G, cfg, batch)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["loss0"]), rtol=1e-5)
    n_checked = 0
    for k in g.files:
        if k.startswith("grad0/"):
            ours = P[k[len("grad0/"):]].grad
            assert ours is not None, k
            ref = g[k]
            scale = max(np.abs(ref).max(), 1e-6)      # exact-zero grads carry only rounding noise
            assert np.abs(ours.numpy() - ref).max() / scale < 2e-5, k
            n_checked += 1
    assert n_checked >= 2


def test_smore_spectrum_matches_reference(tiny_data):
    g = golden("tiny_smore")
    P = _params(g)
    ic, tc, fc = ops.spectrum_convolution(
        torch.from_numpy(g["spec/image_feats"]), torch.from_numpy(g["spec/text_feats"]),
        P["image_complex_weight"], P["text_complex_weight"], P["fusion_complex_weight"], True)
    for ours, key in ((ic, "image_conv"), (tc, "text_conv"), (fc, "fusion_conv")):
        ref = g["spec/" + key]
        assert np.abs(ours.numpy() - ref).max() / np.abs(ref).max() < 1e-5


@pytest.mark.parametrize("model", list(TAGS))
def test_scores_topk_metrics(model):
    g = golden(TAGS[model])
    ue, ie = torch.from_numpy(g["eval_user_emb"]), torch.from_numpy(g["eval_item_emb"])
    users = torch.from_numpy(g["eval_batch_users"])
    mask = torch.from_numpy(g["eval_batch_mask"])
    scores = ops.full_sort_scores(ue, ie, users)
    np.testing.assert_allclose(scores.numpy(), g["eval_scores"], rtol=1e-5, atol=1e-7)
    ours = ops.mask_topk(torch.from_numpy(g["eval_scores"]), mask[0], mask[1], 50).numpy()
    ref = g["eval_topk_ref"]
    # identical unless the reference's arbitrary tie order differs: compare as score sequences
    s = g["eval_scores"].copy()
    s[mask[0].numpy(), mask[1].numpy()] = -1e10
    assert np.array_equal(np.take_along_axis(s, ours, 1), np.take_along_axis(s, ref, 1))
    same = (ours == ref).mean()
    assert same > 0.99


# --------------------------------------------------------------------------- trainer trajectory
def replay_golden_prologue(train_loader):
    """RNG consumption of tests/golden/make_golden.py before its two-epoch loop: seed, item
    shuffle, one dataset shuffle and two batches of negatives."""
    pkg("config").init_seed(999)
    train_loader.pretrain_setup()
    it = iter(train_loader)
    next(it), next(it)
    train_loader.pr = 0


def _loaders():
    import torch as _t
    synth, cfgm, data_m = pkg("synth"), pkg("config"), pkg("data")
    from conftest import TINY
    data = synth.make_dataset("tiny", **TINY)
    config = cfgm.Config("LightGCN", "tiny", {"device": _t.device("cpu"), "data_path": None})
    ds = data_m.RecDataset(config, data.users, data.items, data.labels)
    tr, va, te = ds.split()
    return data_m.TrainDataLoader(config, tr, batch_size=512, shuffle=True), \
        data_m.EvalDataLoader(config, va, additional_dataset=tr, batch_size=64)


@pytest.mark.parametrize("model,tag,mg", [("LayerGCN", "tiny_layergcn", False), ("MGCN", "tiny_mgcn", False),
                                          ("SMORE", "tiny_smore", True), ("SMORE", "tiny_smore_nomg", False)])
def test_oracle_trainer_trajectory(model, tag, mg, tiny_data, tiny_train):
    """Two epochs through the restated training step (incl. the mirror-gradient schedule) land on
    the reference's losses and parameters."""
    from oracle.train import OracleTrainer
    g = golden(tag)
    G, _ = _graphs(model, tiny_data, tiny_train)
    G.setdefault("masked_adj", G["norm_adj"])
    cfg = dict(CFG[model])
    sched = (0.96, 50) if model in ("MGCN", "SMORE") else (1.0, 50)
    tr = OracleTrainer(model, _params(g), G, cfg, lr=1e-3, lr_scheduler=sched, mg_enable=mg)
    train, valid = _loaders()
    replay_golden_prologue(train)
    losses = [tr.train_epoch(train) for _ in range(2)]
    np.testing.assert_allclose(losses, g["fit/train_loss"], rtol=2e-5)
    if "fit/global_step" in g.files:
        assert tr.global_step == int(g["fit/global_step"])
    for k in g.files:
        if k.startswith("fit/param/"):
            ref = g[k]
            ours = tr.P[k[len("fit/param/"):]].detach().numpy()
            assert np.abs(ours - ref).max() / np.abs(ref).max() < 1e-4, k
    # final valid metrics through the oracle's scoring / top-K / metric formulas
    ue, ie = tr.embeddings()
    rows = []
    for b in valid:
        s = ops.full_sort_scores(ue, ie, b[0])
        rows.append(ops.mask_topk(s, b[1][0], b[1][1], 50).numpy())
    raw = ops.calculate_metrics(valid.get_eval_items(), valid.get_eval_len_list(), np.concatenate(rows))
    np.testing.assert_allclose(raw, g["fit/valid_metrics_raw"], atol=2e-3)
