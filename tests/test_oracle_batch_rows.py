"""The algebra behind the product's batch-row training step, pinned on the reference's own math (CPU, float64).

models.SMORE / models.MGCN evaluate the row-local tail of the forward (preference module / attention fuser,
`content + side`) on the 3 B rows `users | n_users + pos | n_users + neg` of the batch only, with the gathers of
smore.py:395-407 / mgcn.py:238-251 hoisted above it (csrc/batch_rows.cu). This file restates that transformation
with the oracle's operators and checks, in float64 where rounding cannot hide a difference, that the loss and
EVERY parameter gradient equal those of the oracle's all-rows loss (oracle/models.py, itself pinned to the
reference's golden vectors by test_oracle_golden.py) -- including batches with repeated users, repeated items
and pos = neg collisions."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from oracle import build, models, ops
from test_oracle_golden import CFG, TAGS

F64 = torch.float64


def _lin(P, name, x):
    return F.linear(x, P[name + ".weight"], P.get(name + ".bias"))


def _setup(model, tiny_data, tiny_train):
    g = golden(TAGS[model])
    u, i = tiny_train
    G, _ = build.build_graphs(model, u, i, tiny_data.n_users, tiny_data.n_items, tiny_data.image_feat,
                              tiny_data.text_feat, CFG[model], dtype=F64)
    P = {k[len("param0/"):]: torch.from_numpy(g[k]).to(F64).requires_grad_(True) for k in g.files
         if k.startswith("param0/")}
    batch = torch.from_numpy(g["batch0"]).clone()
    batch[0, 1::7] = batch[0, 0]                       # one user many times
    batch[1, 2::5] = batch[1, 1]                       # one positive item many times
    batch[2, ::3] = batch[1, ::3].roll(1)              # negatives that are someone's positive
    batch[2, 5] = batch[1, 5]                          # pos = neg
    return g, G, P, batch


def _views(P, G, cfg, xs, keys):
    out = []
    for x, key in zip(xs, keys):
        for _ in range(cfg["n_layers"]):
            x = ops.spmm(G[key], x)
        out.append(torch.cat([ops.spmm(G["R"], x), x], 0))
    return out


def _batch_loss(cfg, all_e, side, content, B, temperature):
    """Compact tables [3 B, d]: rows 0..B-1 users, B..2B-1 positives, 2B..3B-1 negatives."""
    loss = models._reg_bpr(cfg, all_e[:B], all_e[B:2 * B], all_e[2 * B:])
    cl = ops.infonce(side[B:2 * B], content[B:2 * B], temperature) + ops.infonce(side[:B], content[:B], temperature)
    return loss + cfg["cl_loss"] * cl


def mgcn_batch_rows_loss(P, G, cfg, batch):
    """mgcn.py:146-253 with the tail on the batch rows."""
    u, p, n = batch[0], batch[1], batch[2]
    img = _lin(P, "image_trs", P["image_embedding.weight"])
    txt = _lin(P, "text_trs", P["text_embedding.weight"])
    item, user = P["item_id_embedding.weight"], P["user_embedding.weight"]
    nu = user.shape[0]
    img_i = item * torch.sigmoid(_lin(P, "gate_v.0", img))
    txt_i = item * torch.sigmoid(_lin(P, "gate_t.0", txt))
    content = ops.propagate_mean(G["norm_adj"], torch.cat([user, item], 0), cfg["n_ui_layers"])
    img_e, txt_e = _views(P, G, cfg, (img_i, txt_i), ("image_adj", "text_adj"))
    rows = torch.cat([u, nu + p, nu + n])
    img_e, txt_e, content = img_e[rows], txt_e[rows], content[rows]          # <- the hoisted gathers

    def q(x):
        return F.linear(torch.tanh(_lin(P, "query_common.0", x)), P["query_common.2.weight"])
    att = torch.softmax(torch.cat([q(img_e), q(txt_e)], dim=-1), dim=-1)
    common = att[:, 0].unsqueeze(1) * img_e + att[:, 1].unsqueeze(1) * txt_e
    sep_i = torch.sigmoid(_lin(P, "gate_image_prefer.0", content)) * (img_e - common)
    sep_t = torch.sigmoid(_lin(P, "gate_text_prefer.0", content)) * (txt_e - common)
    side = (sep_i + sep_t + common) / 3
    return _batch_loss(cfg, content + side, side, content, u.numel(), 0.2)


def smore_batch_rows_loss(P, G, cfg, batch):
    """smore.py:255-411 with the tail on the batch rows (dropout off)."""
    u, p, n = batch[0], batch[1], batch[2]
    img = _lin(P, "image_trs", P["image_embedding.weight"])
    txt = _lin(P, "text_trs", P["text_embedding.weight"])
    ic, tc, fc = ops.spectrum_convolution(img, txt, P["image_complex_weight"], P["text_complex_weight"],
                                          P["fusion_complex_weight"], cfg.get("spectral_weight_norm", True))
    item, user = P["item_id_embedding.weight"], P["user_embedding.weight"]
    nu = user.shape[0]
    s = cfg.get("inject_scale", 0.7)
    xs = [item + s * torch.sigmoid(_lin(P, name, c)) for name, c in (("gate_v.0", ic), ("gate_t.0", tc), ("gate_f.0", fc))]
    content = ops.propagate_mean(G["norm_adj"], torch.cat([user, item], 0), cfg["n_ui_layers"])
    img_e, txt_e, fus_e = _views(P, G, cfg, xs, ("image_adj", "text_adj", "fusion_adj"))
    rows = torch.cat([u, nu + p, nu + n])
    img_e, txt_e, fus_e, content = img_e[rows], txt_e[rows], fus_e[rows], content[rows]   # <- the hoisted gathers

    def q(name, x):
        return F.linear(torch.tanh(_lin(P, name + ".0", x)), P[name + ".2.weight"])
    agg_i = torch.softmax(q("query_v", fus_e), dim=-1) * img_e
    agg_t = torch.softmax(q("query_t", fus_e), dim=-1) * txt_e
    pi = torch.sigmoid(_lin(P, "gate_image_prefer.0", content))
    pt = torch.sigmoid(_lin(P, "gate_text_prefer.0", content))
    pf = torch.sigmoid(_lin(P, "gate_fusion_prefer.0", content))
    side = torch.mean(torch.stack([pi * agg_i, pt * agg_t, pf * fus_e]), dim=0)
    return _batch_loss(cfg, content + side, side, content, u.numel(), cfg.get("cl_temp", 0.2))


BATCH_ROWS = {"MGCN": mgcn_batch_rows_loss, "SMORE": smore_batch_rows_loss}


@pytest.mark.parametrize("model", ["MGCN", "SMORE"])
def test_batch_rows_step_is_the_all_rows_step(model, tiny_data, tiny_train):
    g, G, P, batch = _setup(model, tiny_data, tiny_train)
    cfg = CFG[model]
    ref = models.LOSS[model](P, G, cfg, batch)
    ref.backward()
    g_ref = {k: v.grad.clone() for k, v in P.items() if v.grad is not None}
    for v in P.values():
        v.grad = None
    ours = BATCH_ROWS[model](P, G, cfg, batch)
    ours.backward()
    assert abs(ours.item() - ref.item()) <= 1e-13 * abs(ref.item())
    assert {k for k, v in P.items() if v.grad is not None} == set(g_ref)
    # float64: what is left is the order of the sums over rows; the floor covers gradients that are themselves a
    # cancelling sum (MGCN's query_common.0.bias: softmax over two logits makes its terms cancel pairwise)
    floor = 1e-13 * max(float(v.abs().max()) for v in g_ref.values())
    for k, v in g_ref.items():
        scale = float(v.abs().max())
        assert float((P[k].grad - v).abs().max()) <= 1e-11 * scale + floor, k


@pytest.mark.parametrize("model", ["MGCN", "SMORE"])
def test_batch_rows_restatement_reproduces_the_reference_loss(model, tiny_data, tiny_train):
    """The same restatement on the reference's own first batch against the loss the unmodified reference printed."""
    g, G, P, _ = _setup(model, tiny_data, tiny_train)
    with torch.no_grad():
        loss = BATCH_ROWS[model](P, G, CFG[model], torch.from_numpy(g["batch0"]))
    np.testing.assert_allclose(loss.item(), float(g["loss0"]), rtol=1e-5)
