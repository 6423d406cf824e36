"""Oracle (test infrastructure): adjacency construction, bit-exact restatements.

All functions return COO parts `(rows int64, cols int64, vals float32)` in the order the
reference produces them, so tests can compare index arrays directly.
"""
from __future__ import annotations

import numpy as np
import torch


def _mirror_sorted(users, items, n_users):
    """Row-major sorted COO pattern of A = [[0, R], [R^T, 0]] (the scipy CSR->COO order of
    layergcn.py:108-113 and mgcn.py:109-136)."""
    users = np.asarray(users, np.int64)
    items = np.asarray(items, np.int64)
    o1 = np.lexsort((items, users))
    o2 = np.lexsort((users, items))
    rows = np.concatenate([users[o1], items[o2] + n_users])
    cols = np.concatenate([items[o1] + n_users, users[o2]])
    return rows, cols


def edge_list(users, items):
    """`dataset.inter_matrix('coo').astype(np.float32)` (layergcn.py:20-21, freedom.py:43): with
    the installed scipy (1.18) `astype` returns the COO in canonical form, i.e. edges sorted by
    (user, item); this is the edge order `get_edge_info` / torch.multinomial index into
    (layergcn.py:83-89). (The pinned scipy 1.7.3 kept file order; the adjacency is the same.)"""
    users = np.asarray(users, np.int64)
    items = np.asarray(items, np.int64)
    o = np.lexsort((items, users))
    return users[o], items[o]


def norm_adj_f64eps(users, items, n_users, n_items):
    """LayerGCN/FREEDOM/LightGCN.get_norm_adj_mat (layergcn.py:91-117, freedom.py:102-128,
    lightgcn.py:65-103): deg = count of non-zeros per row; s = (deg + 1e-7) ** -0.5 in float64;
    value = float32(s_i * s_j) with the product taken in float64 (scipy D * A * D)."""
    n = n_users + n_items
    rows, cols = _mirror_sorted(users, items, n_users)
    deg = np.bincount(rows, minlength=n)
    s = np.power(deg + 1e-7, -0.5)                        # float64
    vals = ((s[rows] * np.float64(1.0)) * s[cols]).astype(np.float32)
    return rows, cols, vals


def norm_adj_f32(users, items, n_users, n_items):
    """MGCN/SMORE.get_adj_mat (mgcn.py:109-136, smore.py:176-199): rowsum float32;
    s = rowsum ** -0.5 in float32 with inf -> 0; value = (s_i * 1) * s_j in float32."""
    n = n_users + n_items
    rows, cols = _mirror_sorted(users, items, n_users)
    rowsum = np.bincount(rows, minlength=n).astype(np.float32)
    with np.errstate(divide="ignore"):
        s = np.power(rowsum, np.float32(-0.5)).astype(np.float32)
    s[np.isinf(s)] = 0.0
    vals = (s[rows] * np.float32(1.0)) * s[cols]
    return rows, cols, vals.astype(np.float32)


def r_block(rows, cols, vals, n_users):
    """R = norm_adj[:U, U:] (mgcn.py:134, smore.py:198)."""
    m = rows < n_users
    return rows[m], cols[m] - n_users, vals[m]


def edge_norm_f32(users, items, n_users, n_items):
    """LayerGCN/FREEDOM._normalize_adj_m (layergcn.py:72-81, freedom.py:147-156): int64 degree
    sums, `1e-7 + deg` promotes to float32, pow -0.5 and the product in float32 (torch CPU)."""
    u = torch.as_tensor(np.asarray(users), dtype=torch.int64)
    i = torch.as_tensor(np.asarray(items), dtype=torch.int64)
    row_sum = 1e-7 + torch.bincount(u, minlength=n_users)
    col_sum = 1e-7 + torch.bincount(i, minlength=n_items)
    r = torch.pow(row_sum, -0.5)
    c = torch.pow(col_sum, -0.5)
    return (r[u] * c[i]).numpy()


def masked_adj(users, items, keep_idx, n_users, n_items):
    """LayerGCN/FREEDOM.pre_epoch_processing after the RNG draw (layergcn.py:63-70,
    freedom.py:138-145): keep edges `keep_idx`, re-normalise, mirror; unsorted COO."""
    ku = np.asarray(users)[keep_idx]
    ki = np.asarray(items)[keep_idx]
    v = edge_norm_f32(ku, ki, n_users, n_items)
    rows = np.concatenate([ku, ki + n_users])
    cols = np.concatenate([ki + n_users, ku])
    return rows, cols, np.concatenate([v, v])


def build_sim(feat: torch.Tensor):
    """utils.py:134-137."""
    n = feat.div(torch.norm(feat, p=2, dim=-1, keepdim=True))
    return torch.mm(n, n.transpose(1, 0))


def freedom_knn_adj(feat: torch.Tensor, k: int):
    """FREEDOM.get_knn_adj_mat + compute_normalized_laplacian (freedom.py:79-100): binary kNN
    edges, value = (k+1e-7)^-1/2 (row) * (rowsum(col)+1e-7)^-1/2, float32."""
    sim = build_sim(feat)
    _, knn_ind = torch.topk(sim, k, dim=-1)
    n = sim.shape[0]
    rows = torch.arange(n).unsqueeze(1).expand(-1, k).flatten()
    cols = knn_ind.flatten()
    row_sum = 1e-7 + torch.bincount(rows, minlength=n)
    r = torch.pow(row_sum, -0.5)
    vals = r[rows] * r[cols]
    return rows.numpy(), cols.numpy(), vals.numpy()


def freedom_mm_adj(image_feat, text_feat, k, w_image):
    """freedom.py:68-77: w * image_adj + (1-w) * text_adj, coalesced (duplicates summed)."""
    ri, ci, vi = freedom_knn_adj(image_feat, k)
    rt, ct, vt = freedom_knn_adj(text_feat, k)
    n = image_feat.shape[0]
    a = torch.sparse_coo_tensor(np.vstack([ri, ci]), torch.from_numpy(vi), (n, n))
    b = torch.sparse_coo_tensor(np.vstack([rt, ct]), torch.from_numpy(vt), (n, n))
    m = (w_image * a + (1.0 - w_image) * b).coalesce()
    idx = m.indices().numpy()
    return idx[0], idx[1], m.values().numpy()


def knn_sym_graph(feat: torch.Tensor, k: int):
    """build_sim + build_knn_normalized_graph(norm_type='sym') + get_sparse_laplacian
    (utils.py:134-152, 171-184): weighted kNN edges, deg = sum of weights per *row*,
    w <- deg_r^-1/2 * w * deg_c^-1/2 (inf -> 0), float32. Not symmetric."""
    sim = build_sim(feat)
    knn_val, knn_ind = torch.topk(sim, k, dim=-1)
    n = sim.shape[0]
    rows = torch.arange(n).unsqueeze(1).expand(-1, k).flatten()
    cols = knn_ind.flatten()
    w = knn_val.flatten()
    deg = torch.zeros(n, dtype=w.dtype).index_add_(0, rows, w)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    w = dis[rows] * w * dis[cols]
    return rows.numpy(), cols.numpy(), w.numpy()


def max_pool_fusion(a, b, n):
    """SMORE.max_pool_fusion (smore.py:153-174): union of two edge sets, element-wise max,
    row-major sorted."""
    ra, ca, va = a
    rb, cb, vb = b
    ka = np.asarray(ra, np.int64) * n + ca
    kb = np.asarray(rb, np.int64) * n + cb
    keys = np.union1d(ka, kb)
    va_full = np.full(len(keys), -np.inf, dtype=np.float32)
    vb_full = np.full(len(keys), -np.inf, dtype=np.float32)
    va_full[np.searchsorted(keys, ka)] = va
    vb_full[np.searchsorted(keys, kb)] = vb
    return keys // n, keys % n, np.maximum(va_full, vb_full)


def to_torch_csr(rows, cols, vals, shape, dtype=torch.float32):
    """Coalesced CSR tensor for the oracle's SpMM (sums duplicates like torch.sparse.mm does)."""
    t = torch.sparse_coo_tensor(np.vstack([rows, cols]), torch.as_tensor(vals).to(dtype), shape)
    return t.coalesce().to_sparse_csr()
