"""Oracle (test infrastructure): assemble the graph dict `G` each model needs from raw edges."""
from __future__ import annotations

import numpy as np
import torch

from . import graph


def build_graphs(model, users, items, n_users, n_items, image_feat=None, text_feat=None, cfg=None,
                 dtype=torch.float32, keep_idx=None):
    """Returns (G, parts): G maps names to torch CSR tensors, parts to the raw COO triples."""
    cfg = cfg or {}
    n = n_users + n_items
    parts = {}
    if model in ("LightGCN", "LayerGCN", "FREEDOM"):
        parts["norm_adj"] = graph.norm_adj_f64eps(users, items, n_users, n_items)
        if keep_idx is not None:
            parts["masked_adj"] = graph.masked_adj(users, items, keep_idx, n_users, n_items)
        else:
            parts["masked_adj"] = parts["norm_adj"]
    else:
        parts["norm_adj"] = graph.norm_adj_f32(users, items, n_users, n_items)
        parts["R"] = graph.r_block(*parts["norm_adj"], n_users)
    if model == "FREEDOM":
        parts["mm_adj"] = graph.freedom_mm_adj(torch.as_tensor(image_feat), torch.as_tensor(text_feat),
                                               cfg["knn_k"], cfg["mm_image_weight"])
    if model == "MGCN":
        parts["image_adj"] = graph.knn_sym_graph(torch.as_tensor(image_feat), cfg["knn_k"])
        parts["text_adj"] = graph.knn_sym_graph(torch.as_tensor(text_feat), cfg["knn_k"])
    if model == "SMORE":
        parts["image_adj"] = graph.knn_sym_graph(torch.as_tensor(image_feat), cfg["image_knn_k"])
        parts["text_adj"] = graph.knn_sym_graph(torch.as_tensor(text_feat), cfg["text_knn_k"])
        parts["fusion_adj"] = graph.max_pool_fusion(parts["image_adj"], parts["text_adj"], n_items)
    shapes = {"norm_adj": (n, n), "masked_adj": (n, n), "R": (n_users, n_items)}
    G = {k: graph.to_torch_csr(*v, shapes.get(k, (n_items, n_items)), dtype) for k, v in parts.items()}
    return G, parts
