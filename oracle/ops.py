"""Oracle (test infrastructure): the hot-path operators restated on torch CPU ops.

Every function works in whatever floating dtype its inputs carry (float32 to mimic the reference,
float64 as the high-precision yardstick) and is differentiable through autograd.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def spmm(A, X):
    """torch.sparse.mm(A, X) -- layergcn.py:133, freedom.py:169,174, mgcn.py:162-184,
    smore.py:282-317, lightgcn.py:122."""
    return torch.sparse.mm(A, X)


def propagate_mean(A, X0, n_layers):
    """LightGCN-style propagation: mean over layers 0..L (freedom.py:171-179, mgcn.py:159-167,
    smore.py:278-287, lightgcn.py:118-128)."""
    e, layers = X0, [X0]
    for _ in range(n_layers):
        e = spmm(A, e)
        layers.append(e)
    return torch.stack(layers, dim=1).mean(dim=1)


def layergcn_propagate(A, X0, n_layers):
    """LayerGCN.forward (layergcn.py:127-140): each layer is re-weighted by its cosine similarity
    to the ego layer (torch 2.x semantics: each vector divided by max(norm, 1e-8)), the
    re-weighted layer feeds the next one, output = sum of layers 1..L."""
    e, layers = X0, []
    for _ in range(n_layers):
        e = spmm(A, e)
        w = F.cosine_similarity(e, X0, dim=-1)
        e = torch.einsum("a,ab->ab", w, e)
        layers.append(e)
    return torch.sum(torch.stack(layers, dim=0), dim=0)


def unit_mag(wc):
    """smore.py:221-229."""
    return wc / (torch.abs(wc) + 1e-8)


def spectrum_convolution(image, text, w_img, w_txt, w_fus, weight_norm=True):
    """SMORE.spectrum_convolution (smore.py:209-238). Weights are real [1, d/2+1, 2]."""
    d = image.shape[1]
    fi = torch.fft.rfft(image, dim=1, norm="ortho")
    ft = torch.fft.rfft(text, dim=1, norm="ortho")
    wi, wt, wf = (torch.view_as_complex(w.contiguous()) for w in (w_img, w_txt, w_fus))
    if weight_norm:
        wi, wt, wf = unit_mag(wi), unit_mag(wt), unit_mag(wf)
    ic = torch.fft.irfft(fi * wi, n=d, dim=1, norm="ortho")
    tc = torch.fft.irfft(ft * wt, n=d, dim=1, norm="ortho")
    fc = torch.fft.irfft(ft * fi * wf, n=d, dim=1, norm="ortho")
    return ic, tc, fc


def bpr_scores(u, p, n):
    return torch.mul(u, p).sum(dim=1), torch.mul(u, n).sum(dim=1)


def bpr_sum(u, p, n):
    """LayerGCN.bpr_loss (layergcn.py:142-154): sum of -logsigmoid."""
    ps, ns = bpr_scores(u, p, n)
    return torch.sum(-F.logsigmoid(ps - ns))


def bpr_mean(u, p, n):
    """FREEDOM/MGCN/SMORE.bpr_loss (freedom.py:182-189, mgcn.py:210-222, smore.py:366-378)."""
    ps, ns = bpr_scores(u, p, n)
    return -torch.mean(F.logsigmoid(ps - ns))


def bpr_gamma_mean(u, p, n, gamma=1e-10):
    """BPRLoss used by LightGCN (loss.py:28-36)."""
    ps, ns = bpr_scores(u, p, n)
    return -torch.log(gamma + torch.sigmoid(ps - ns)).mean()


def l2_half(*embs):
    """L2Loss (loss.py:54-61) and the MGCN/SMORE regulariser numerator: sum of 0.5*||e||^2."""
    return sum(0.5 * torch.sum(e ** 2) for e in embs)


def emb_loss(*embs):
    """EmbLoss (loss.py:39-51): sum of L2 norms / batch."""
    return sum(torch.norm(e, p=2) for e in embs) / embs[-1].shape[0]


def infonce(v1, v2, temperature):
    """MGCN/SMORE.InfoNCE (mgcn.py:224-231, smore.py:380-387)."""
    v1, v2 = F.normalize(v1, dim=1), F.normalize(v2, dim=1)
    pos = torch.exp((v1 * v2).sum(dim=-1) / temperature)
    ttl = torch.exp(torch.matmul(v1, v2.transpose(0, 1)) / temperature).sum(dim=1)
    return torch.mean(-torch.log(pos / ttl))


def full_sort_scores(user_e, item_e, users):
    """full_sort_predict tail (layergcn.py:185-188 etc.)."""
    return torch.matmul(user_e[users], item_e.transpose(0, 1))


def mask_topk(scores, mask_rows, mask_cols, k):
    """Trainer.evaluate (trainer.py:522-526) with the deterministic tie rule 'lower id first':
    stable descending sort. The reference's torch.topk has arbitrary tie order (SURVEY A)."""
    s = scores.clone()
    s[mask_rows, mask_cols] = -1e10
    return torch.sort(s, dim=-1, descending=True, stable=True)[1][:, :k]


# ---------------------------------------------------------------------------------------------
# metrics: topk_evaluator.py:88-101 + metrics.py:12-109 (numpy float64)
def hit_matrix(pos_items, topk_index):
    return np.asarray([np.isin(row, gt) for gt, row in zip(pos_items, topk_index)])


def recall_(hits, pos_len):
    return (np.cumsum(hits, axis=1) / pos_len.reshape(-1, 1)).mean(axis=0)


def precision_(hits, pos_len):
    return (hits.cumsum(axis=1) / np.arange(1, hits.shape[1] + 1)).mean(axis=0)


def ndcg_(hits, pos_len):
    k = hits.shape[1]
    disc = 1.0 / np.log2(np.arange(1, k + 1, dtype=np.float64) + 1)
    idcg_all = np.cumsum(disc)
    idcg_len = np.minimum(pos_len, k)
    pos = np.minimum(np.arange(k)[None, :], idcg_len[:, None] - 1)
    idcg = idcg_all[pos]
    dcg = np.cumsum(np.where(hits, disc[None, :], 0.0), axis=1)
    return (dcg / idcg).mean(axis=0)


def map_(hits, pos_len):
    k = hits.shape[1]
    pre = hits.cumsum(axis=1) / np.arange(1, k + 1)
    sum_pre = np.cumsum(pre * hits.astype(np.float64), axis=1)
    actual = np.minimum(pos_len, k)
    ranges = np.minimum(np.arange(1, k + 1)[None, :], actual[:, None])
    return (sum_pre / ranges).mean(axis=0)


METRICS = {"recall": recall_, "ndcg": ndcg_, "precision": precision_, "map": map_}


def calculate_metrics(pos_items, pos_len, topk_index, names=("recall", "ndcg", "precision", "map")):
    hits = hit_matrix(pos_items, topk_index)
    return np.stack([METRICS[n](hits, np.asarray(pos_len)) for n in names], axis=0)
