"""Oracle (test infrastructure): the four models (+LightGCN) as pure functions of a parameter
dict `P` keyed by the reference's `state_dict` names and a graph dict `G` of torch sparse CSR
tensors. torch CPU, autograd gives the gradients.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def _lin(P, name, x):
    b = P.get(name + ".bias")
    return F.linear(x, P[name + ".weight"], b)


# ----------------------------------------------------------------------------- LightGCN
def lightgcn_forward(P, G, cfg):
    """lightgcn.py:118-131."""
    x0 = torch.cat([P["embedding_dict.user_emb"], P["embedding_dict.item_emb"]], 0)
    out = ops.propagate_mean(G["norm_adj"], x0, cfg["n_layers"])
    nu = P["embedding_dict.user_emb"].shape[0]
    return out[:nu], out[nu:]


def lightgcn_loss(P, G, cfg, batch):
    """lightgcn.py:133-159."""
    u, p, n = batch[0], batch[1], batch[2]
    ue, ie = lightgcn_forward(P, G, cfg)
    mf = ops.bpr_gamma_mean(ue[u], ie[p], ie[n])
    reg = ops.emb_loss(P["embedding_dict.user_emb"][u], P["embedding_dict.item_emb"][p],
                       P["embedding_dict.item_emb"][n])
    return mf + cfg["reg_weight"] * reg


# ----------------------------------------------------------------------------- LayerGCN
def layergcn_forward(P, G, cfg, adj_key="norm_adj"):
    """layergcn.py:127-140."""
    x0 = torch.cat([P["user_embeddings"], P["item_embeddings"]], 0)
    out = ops.layergcn_propagate(G[adj_key], x0, cfg["n_layers"])
    nu = P["user_embeddings"].shape[0]
    return out[:nu], out[nu:]


def layergcn_loss(P, G, cfg, batch):
    """layergcn.py:165-177 (masked adjacency in training)."""
    u, p, n = batch[0], batch[1], batch[2]
    ue, ie = layergcn_forward(P, G, cfg, "masked_adj")
    mf = ops.bpr_sum(ue[u], ie[p], ie[n])
    reg = ops.l2_half(P["user_embeddings"][u], P["item_embeddings"][p], P["item_embeddings"][n])
    return mf + cfg["reg_weight"] * reg


# ----------------------------------------------------------------------------- FREEDOM
def freedom_forward(P, G, cfg, adj_key="norm_adj"):
    """freedom.py:166-180."""
    h = P["item_id_embedding.weight"]
    for _ in range(cfg["n_mm_layers"]):
        h = ops.spmm(G["mm_adj"], h)
    x0 = torch.cat([P["user_embedding.weight"], P["item_id_embedding.weight"]], 0)
    out = ops.propagate_mean(G[adj_key], x0, cfg["n_ui_layers"])
    nu = P["user_embedding.weight"].shape[0]
    return out[:nu], out[nu:] + h


def freedom_loss(P, G, cfg, batch):
    """freedom.py:191-212."""
    u, p, n = batch[0], batch[1], batch[2]
    ue, ie = freedom_forward(P, G, cfg, "masked_adj")
    loss = ops.bpr_mean(ue[u], ie[p], ie[n])
    txt = _lin(P, "text_trs", P["text_embedding.weight"])
    mf_t = ops.bpr_mean(ue[u], txt[p], txt[n])
    img = _lin(P, "image_trs", P["image_embedding.weight"])
    mf_v = ops.bpr_mean(ue[u], img[p], img[n])
    return loss + cfg["reg_weight"] * (mf_t + mf_v)


# ----------------------------------------------------------------------------- MGCN
def mgcn_forward(P, G, cfg, train=False):
    """mgcn.py:146-208."""
    img = _lin(P, "image_trs", P["image_embedding.weight"])
    txt = _lin(P, "text_trs", P["text_embedding.weight"])
    item = P["item_id_embedding.weight"]
    user = P["user_embedding.weight"]
    nu = user.shape[0]
    img_i = item * torch.sigmoid(_lin(P, "gate_v.0", img))
    txt_i = item * torch.sigmoid(_lin(P, "gate_t.0", txt))
    content = ops.propagate_mean(G["norm_adj"], torch.cat([user, item], 0), cfg["n_ui_layers"])
    for _ in range(cfg["n_layers"]):
        img_i = ops.spmm(G["image_adj"], img_i)
    img_e = torch.cat([ops.spmm(G["R"], img_i), img_i], 0)
    for _ in range(cfg["n_layers"]):
        txt_i = ops.spmm(G["text_adj"], txt_i)
    txt_e = torch.cat([ops.spmm(G["R"], txt_i), txt_i], 0)

    def q(x):
        return F.linear(torch.tanh(_lin(P, "query_common.0", x)), P["query_common.2.weight"])
    att = torch.softmax(torch.cat([q(img_e), q(txt_e)], dim=-1), dim=-1)
    common = att[:, 0].unsqueeze(1) * img_e + att[:, 1].unsqueeze(1) * txt_e
    sep_i = torch.sigmoid(_lin(P, "gate_image_prefer.0", content)) * (img_e - common)
    sep_t = torch.sigmoid(_lin(P, "gate_text_prefer.0", content)) * (txt_e - common)
    side = (sep_i + sep_t + common) / 3
    all_e = content + side
    if train:
        return all_e[:nu], all_e[nu:], side, content
    return all_e[:nu], all_e[nu:]


def _reg_bpr(cfg, u, p, n):
    """mgcn.py:210-222, smore.py:366-378: divides by the *configured* train_batch_size."""
    mf = ops.bpr_mean(u, p, n)
    return mf + cfg["reg_weight"] * (ops.l2_half(u, p, n) / cfg["train_batch_size"])


def mgcn_loss(P, G, cfg, batch):
    """mgcn.py:233-253."""
    u, p, n = batch[0], batch[1], batch[2]
    ue, ie, side, content = mgcn_forward(P, G, cfg, train=True)
    nu = ue.shape[0]
    loss = _reg_bpr(cfg, ue[u], ie[p], ie[n])
    cl = ops.infonce(side[nu:][p], content[nu:][p], 0.2) + ops.infonce(side[:nu][u], content[:nu][u], 0.2)
    return loss + cfg["cl_loss"] * cl


# ----------------------------------------------------------------------------- SMORE
def smore_forward(P, G, cfg, train=False, dropout=None):
    """smore.py:255-364. `dropout` is a callable applied to the three preference gates
    (nn.Dropout in the reference; identity in eval mode)."""
    drop = dropout or (lambda x: x)
    img = _lin(P, "image_trs", P["image_embedding.weight"])
    txt = _lin(P, "text_trs", P["text_embedding.weight"])
    ic, tc, fc = ops.spectrum_convolution(
        img, txt, P["image_complex_weight"], P["text_complex_weight"], P["fusion_complex_weight"],
        cfg.get("spectral_weight_norm", True))
    item = P["item_id_embedding.weight"]
    user = P["user_embedding.weight"]
    nu = user.shape[0]
    gv = torch.sigmoid(_lin(P, "gate_v.0", ic))
    gt = torch.sigmoid(_lin(P, "gate_t.0", tc))
    gf = torch.sigmoid(_lin(P, "gate_f.0", fc))
    if cfg.get("inject_mode", "residual") == "mul":
        img_i, txt_i, fus_i = item * gv, item * gt, item * gf
    else:
        s = cfg.get("inject_scale", 0.7)
        img_i, txt_i, fus_i = item + s * gv, item + s * gt, item + s * gf
    content = ops.propagate_mean(G["norm_adj"], torch.cat([user, item], 0), cfg["n_ui_layers"])
    views = []
    for x, key in ((img_i, "image_adj"), (txt_i, "text_adj"), (fus_i, "fusion_adj")):
        for _ in range(cfg["n_layers"]):
            x = ops.spmm(G[key], x)
        views.append(torch.cat([ops.spmm(G["R"], x), x], 0))
    img_e, txt_e, fus_e = views

    def q(name, x):
        return F.linear(torch.tanh(_lin(P, name + ".0", x)), P[name + ".2.weight"])
    agg_i = torch.softmax(q("query_v", fus_e), dim=-1) * img_e
    agg_t = torch.softmax(q("query_t", fus_e), dim=-1) * txt_e
    pi = drop(torch.sigmoid(_lin(P, "gate_image_prefer.0", content)))
    pt = drop(torch.sigmoid(_lin(P, "gate_text_prefer.0", content)))
    pf = drop(torch.sigmoid(_lin(P, "gate_fusion_prefer.0", content)))
    side = torch.mean(torch.stack([pi * agg_i, pt * agg_t, pf * fus_e]), dim=0)
    all_e = content + side
    if train:
        return all_e[:nu], all_e[nu:], side, content
    return all_e[:nu], all_e[nu:]


def smore_loss(P, G, cfg, batch, dropout=None):
    """smore.py:389-411."""
    u, p, n = batch[0], batch[1], batch[2]
    ue, ie, side, content = smore_forward(P, G, cfg, train=True, dropout=dropout)
    nu = ue.shape[0]
    loss = _reg_bpr(cfg, ue[u], ie[p], ie[n])
    t = cfg.get("cl_temp", 0.2)
    cl = ops.infonce(side[nu:][p], content[nu:][p], t) + ops.infonce(side[:nu][u], content[:nu][u], t)
    return loss + cfg["cl_loss"] * cl


FORWARD = {"LightGCN": lightgcn_forward, "LayerGCN": layergcn_forward, "FREEDOM": freedom_forward,
           "MGCN": mgcn_forward, "SMORE": smore_forward}
LOSS = {"LightGCN": lightgcn_loss, "LayerGCN": layergcn_loss, "FREEDOM": freedom_loss,
        "MGCN": mgcn_loss, "SMORE": smore_loss}
