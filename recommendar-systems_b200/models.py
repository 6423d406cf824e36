"""LayerGCN, LightGCN, FREEDOM, MGCN and SMORE on the B200 operator layer.

Same constructor `(config, dataloader)`, parameter names (state_dicts are interchangeable with the
reference), RNG consumption order at init (same seed -> same initial parameters) and methods
(`pre_epoch_processing`, `forward`, `calculate_loss`, `full_sort_predict`) as
/root/reference/src/models/{layergcn,lightgcn,freedom,mgcn,smore}.py; the torch.sparse.mm /
loss / spectral call sites are replaced by `ops.*` (hand-written sm_100a kernels). Adjacency
tensors are `graph.CSRGraph` objects instead of torch sparse COO tensors.

Additions that the reference does not have: `restore_embeddings()` (one propagation shared by
all eval batches; parameters are frozen under eval) and `full_sort_topk()` (fused
score + mask + top-K).
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import graph as G
from . import ops


class AbstractRecommender(nn.Module):
    """common/abstract_recommender.py:10-67."""

    def pre_epoch_processing(self):
        pass

    def post_epoch_processing(self):
        pass

    def calculate_loss(self, interaction):
        raise NotImplementedError

    def predict(self, interaction):
        raise NotImplementedError

    def full_sort_predict(self, interaction):
        raise NotImplementedError

    def __str__(self):
        params = sum(int(np.prod(p.size())) for p in self.parameters())
        return super().__str__() + "\nTrainable parameters: {}".format(params)


class GeneralRecommender(AbstractRecommender):
    """common/abstract_recommender.py:70-103. Features come from `config['v_feat']`/`['t_feat']`
    (synthetic tensors) or from the reference's .npy files under data_path/dataset."""

    def __init__(self, config, dataloader):
        super().__init__()
        self.USER_ID, self.ITEM_ID = config["USER_ID_FIELD"], config["ITEM_ID_FIELD"]
        self.NEG_ITEM_ID = config["NEG_PREFIX"] + self.ITEM_ID
        self.n_users = dataloader.dataset.get_user_num()
        self.n_items = dataloader.dataset.get_item_num()
        self.batch_size = config["train_batch_size"]
        self.device = config["device"]
        self.v_feat, self.t_feat = None, None
        if not config["end2end"] and config["is_multimodal_model"]:
            v, t = config["v_feat"], config["t_feat"]
            if v is None and t is None and config["data_path"]:
                path = os.path.abspath(config["data_path"] + config["dataset"])
                vp = os.path.join(path, config["vision_feature_file"])
                tp = os.path.join(path, config["text_feature_file"])
                v = np.load(vp, allow_pickle=True) if os.path.isfile(vp) else None
                t = np.load(tp, allow_pickle=True) if os.path.isfile(tp) else None
            if v is not None:
                self.v_feat = torch.as_tensor(v).type(torch.FloatTensor).to(self.device)
            if t is not None:
                self.t_feat = torch.as_tensor(t).type(torch.FloatTensor).to(self.device)
            assert self.v_feat is not None or self.t_feat is not None, "Features all NONE"
        # training edges in the reference's order: inter_matrix('coo').astype(float32) is in
        # canonical (user, item)-sorted order with the installed scipy (layergcn.py:20-21)
        ds = dataloader.dataset
        order = np.lexsort((ds.items, ds.users))
        self._edge_u = torch.from_numpy(ds.users[order]).to(self.device)
        self._edge_i = torch.from_numpy(ds.items[order]).to(self.device)
        key = ds.users.astype(np.int64) * self.n_items + ds.items
        if len(np.unique(key)) != len(key):
            raise ValueError("duplicate (user, item) training interactions are not supported")
        self._eval_cache = None
        # item-range sharding of the trainable feature tables over a process group
        # (parallel.ShardedRows): config['table_shard'] = (rank, world[, group])
        self._table_rows, self._table_group = None, None
        ts = config.get("table_shard") if hasattr(config, "get") else None
        if ts is not None:
            from . import parallel
            self._table_rows = parallel.ShardedRows(self.n_items, int(ts[0]), int(ts[1]))
            self._table_group = ts[2] if len(ts) > 2 else None

    def _feature_table(self, feat):
        """`nn.Embedding.from_pretrained(feat, freeze=False)` (smore.py:76-77, mgcn.py:62-72,
        freedom.py:48-55); with table sharding only this rank's item rows become a parameter."""
        if self._table_rows is None:
            return nn.Embedding.from_pretrained(feat, freeze=False)
        emb = nn.Embedding.from_pretrained(self._table_rows.local(feat).clone(), freeze=False)
        emb.weight._mmrec_sharded = True
        emb.weight._mmrec_global_numel = int(feat.shape[0]) * int(feat.shape[1])
        return emb

    def _project(self, emb, trs):
        """`trs(emb.weight)`; sharded tables: local rows projected, slices all-gathered."""
        if self._table_rows is None:
            # training through the fused optimizer: the [I, F] gradient stays the rank-d product
            # dY W (ops.LowRankGrad); the Trainer switches this on when its optimizer consumes it
            if (getattr(self, "lowrank_table_grad", False) and torch.is_grad_enabled() and
                    emb.weight.requires_grad and ops.table_lowrank_supported(emb.weight, trs.weight)):
                return ops.table_project(emb, trs.weight, trs.bias)
            return trs(emb.weight)
        from . import parallel
        return parallel.sharded_projection(emb.weight, trs.weight, trs.bias, self._table_rows,
                                           self._table_group)

    def train(self, mode=True):
        self._eval_cache = None
        return super().train(mode)

    def _set_masked_adj(self, g):
        """Replace the per-epoch dropout adjacency (layergcn.py:70, freedom.py:145). A training
        step captured in a CUDA graph has the old CSR's pointers, task count and grid baked in, so
        the adjacency version the trainer keys its graphs on moves with every replacement; the
        previous CSR is kept alive until the next replacement so that a graph captured on it can
        never outlive its buffers."""
        self._retired_adj = self.__dict__.get("masked_adj")
        self.masked_adj = g
        self.graph_version = int(self.__dict__.get("graph_version", 0)) + 1

    # -- evaluation helpers ----------------------------------------------------------------
    def _eval_forward(self):
        raise NotImplementedError

    @torch.no_grad()
    def restore_embeddings(self):
        """(user_e, item_e) of the evaluation forward, computed once per eval pass."""
        if self.training or self._eval_cache is None:
            ue, ie = self._eval_forward()
            if self.training:
                return ue, ie
            self._eval_cache = (ue.contiguous(), ie.contiguous())
        return self._eval_cache

    @torch.no_grad()
    def full_sort_predict(self, interaction):
        """[Bu, n_items] scores (layergcn.py:179-188, freedom.py:214-222, mgcn.py:255-263,
        smore.py:414-422): API-compatible dense path. The trainer uses full_sort_topk."""
        ue, ie = self._eval_forward()
        return ops.score_matrix(ue[interaction[0]], ie)

    @torch.no_grad()
    def full_sort_topk(self, users, k, mask_rowptr=None, mask_cols=None):
        ue, ie = self.restore_embeddings()
        return ops.score_mask_topk(ue, users, ie, k, mask_rowptr, mask_cols)


def _split(x, n_users):
    return x[:n_users], x[n_users:]


# ============================================================================== LightGCN
class LightGCN(GeneralRecommender):
    """models/lightgcn.py (propagation + scoring are the hot path of config 5)."""

    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.latent_dim = config["embedding_size"]
        self.n_layers = config["n_layers"]
        self.reg_weight = config["reg_weight"]
        init = nn.init.xavier_uniform_
        self.embedding_dict = nn.ParameterDict({
            "user_emb": nn.Parameter(init(torch.empty(self.n_users, self.latent_dim))),
            "item_emb": nn.Parameter(init(torch.empty(self.n_items, self.latent_dim)))})
        self.norm_adj_matrix = G.build_ui_graph(self._edge_u, self._edge_i, self.n_users,
                                                self.n_items, "f64eps")

    def get_ego_embeddings(self):
        return torch.cat([self.embedding_dict["user_emb"], self.embedding_dict["item_emb"]], 0)

    def forward(self):
        out = ops.propagate_mean(self.norm_adj_matrix, self.get_ego_embeddings(), self.n_layers)
        return _split(out, self.n_users)

    _eval_forward = forward

    def calculate_loss(self, interaction):
        user, pos, neg = interaction[0], interaction[1], interaction[2]
        ue, ie = self.forward()
        # BPRLoss with gamma (common/loss.py:28-36) and EmbLoss: plain torch, not a named hot op
        ps = (ue[user] * ie[pos]).sum(1)
        ns = (ue[user] * ie[neg]).sum(1)
        mf = -torch.log(1e-10 + torch.sigmoid(ps - ns)).mean()
        e = self.embedding_dict
        reg = (torch.norm(e["user_emb"][user]) + torch.norm(e["item_emb"][pos]) +
               torch.norm(e["item_emb"][neg])) / user.shape[0]
        return mf + self.reg_weight * reg


# ============================================================================== LayerGCN
class LayerGCN(GeneralRecommender):
    """models/layergcn.py:15-188."""

    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.latent_dim = config["embedding_size"]
        self.n_layers = config["n_layers"]
        self.reg_weight = config["reg_weight"]
        self.dropout = config["dropout"]
        self.n_nodes = self.n_users + self.n_items
        self.user_embeddings = nn.Parameter(nn.init.xavier_uniform_(torch.empty(self.n_users, self.latent_dim)))
        self.item_embeddings = nn.Parameter(nn.init.xavier_uniform_(torch.empty(self.n_items, self.latent_dim)))
        self.norm_adj_matrix = G.build_ui_graph(self._edge_u, self._edge_i, self.n_users,
                                                self.n_items, "f64eps")
        self.masked_adj = None
        self.forward_adj = None
        self.pruning_random = False
        # layergcn.py:42,83-89: edge values stay on the CPU (the multinomial draws from the CPU RNG)
        self.edge_values = _edge_values_cpu(self._edge_u, self._edge_i, self.n_users, self.n_items)

    def pre_epoch_processing(self):
        """layergcn.py:51-70: same RNG calls (torch.multinomial on CPU / random.sample,
        alternating); re-normalisation + CSR rebuild on device (K12)."""
        if self.dropout <= .0:
            self.masked_adj = self.norm_adj_matrix
            return
        keep_len = int(self.edge_values.size(0) * (1. - self.dropout))
        if self.pruning_random:
            keep_idx = torch.tensor(random.sample(range(self.edge_values.size(0)), keep_len))
        else:
            keep_idx = torch.multinomial(self.edge_values, keep_len)
        self.pruning_random = True ^ self.pruning_random
        self._set_masked_adj(self._masked_graph(keep_idx))

    def _masked_graph(self, keep_idx):
        keep_idx = keep_idx.to(self.device)
        return G.build_ui_graph(self._edge_u[keep_idx], self._edge_i[keep_idx], self.n_users,
                                self.n_items, "edge_f32")

    def get_ego_embeddings(self):
        return torch.cat([self.user_embeddings, self.item_embeddings], 0)

    def forward(self):
        out = ops.layergcn_propagate(self.forward_adj, self.get_ego_embeddings(), self.n_layers)
        return _split(out, self.n_users)

    def _eval_forward(self):
        self.forward_adj = self.norm_adj_matrix
        return self.forward()

    def calculate_loss(self, interaction):
        """layergcn.py:165-177: sum-BPR on propagated rows + reg_weight * L2 on the ego rows."""
        user, pos, neg = interaction[0], interaction[1], interaction[2]
        self.forward_adj = self.masked_adj
        ego = self.get_ego_embeddings()
        out = ops.layergcn_propagate(self.forward_adj, ego, self.n_layers)
        mf_loss = ops.bpr_table(out, self.n_users, user, pos, neg)[0]
        reg_loss = ops.bpr_table(ego, self.n_users, user, pos, neg)[1]
        return mf_loss + self.reg_weight * reg_loss


def _edge_values_cpu(edge_u, edge_i, n_users, n_items):
    """get_edge_info/_normalize_adj_m (layergcn.py:72-89) on the CPU, float32."""
    u, i = edge_u.cpu(), edge_i.cpu()
    r = torch.pow(1e-7 + torch.bincount(u, minlength=n_users), -0.5)
    c = torch.pow(1e-7 + torch.bincount(i, minlength=n_items), -0.5)
    return r[u] * c[i]


# ----------------------------------------------------------------------------- item graphs
def knn_sym_coo(feat, k, neighbors=None):
    """build_sim + build_knn_normalized_graph(sparse, 'sym') + get_sparse_laplacian
    (utils/utils.py:134-152, 171-184) on the library's kNN kernels (no per-element Python loop,
    no [I, I] torch.topk)."""
    return ops.knn_graph(feat, k, "sym", neighbors)


def freedom_knn_coo(feat, k, neighbors=None):
    """FREEDOM.get_knn_adj_mat + compute_normalized_laplacian (freedom.py:79-100)."""
    return ops.knn_graph(feat, k, "freedom", neighbors)


def _knn_override(config, name):
    """Tests pin the kNN edge sets through config['item_knn'][name] = [I, k] neighbour ids."""
    return (config["item_knn"] or {}).get(name)


def max_pool_fusion_coo(a, b, n):
    """SMORE.max_pool_fusion (smore.py:153-174)."""
    ka = a[0] * n + a[1]
    kb = b[0] * n + b[1]
    keys, inv = torch.unique(torch.cat([ka, kb]), return_inverse=True)
    va = torch.full((keys.numel(),), float("-inf"), device=keys.device)
    vb = torch.full((keys.numel(),), float("-inf"), device=keys.device)
    va[inv[:ka.numel()]] = a[2]
    vb[inv[ka.numel():]] = b[2]
    return keys // n, keys % n, torch.maximum(va, vb)


def _coo_override(config, name, device):
    """Tests inject reference-built graphs through config['item_graphs'][name] = (r, c, v)."""
    g = (config["item_graphs"] or {}).get(name)
    if g is None:
        return None
    return tuple(torch.as_tensor(x).to(device) for x in g)


# ============================================================================== FREEDOM
class FREEDOM(GeneralRecommender):
    """models/freedom.py:22-222."""

    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.embedding_dim = config["embedding_size"]
        self.feat_embed_dim = config["feat_embed_dim"]
        self.knn_k = config["knn_k"]
        self.n_layers = config["n_mm_layers"]
        self.n_ui_layers = config["n_ui_layers"]
        self.reg_weight = config["reg_weight"]
        self.mm_image_weight = config["mm_image_weight"]
        self.dropout = config["dropout"]
        self.n_nodes = self.n_users + self.n_items
        self.norm_adj = G.build_ui_graph(self._edge_u, self._edge_i, self.n_users, self.n_items,
                                         "f64eps")
        self.masked_adj, self.mm_adj = None, None
        # freedom.py:45-46: edge values live on the device (multinomial draws from the CUDA RNG)
        self.edge_values = _edge_values_cpu(self._edge_u, self._edge_i, self.n_users,
                                            self.n_items).to(self.device)
        self.edge_dropout_rng = str(config.get("edge_dropout_rng", "device"))
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_dim)
        self.item_id_embedding = nn.Embedding(self.n_items, self.embedding_dim)
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.item_id_embedding.weight)
        if self.v_feat is not None:
            self.image_embedding = self._feature_table(self.v_feat)
            self.image_trs = ops.Linear(self.v_feat.shape[1], self.feat_embed_dim)
        if self.t_feat is not None:
            self.text_embedding = self._feature_table(self.t_feat)
            self.text_trs = ops.Linear(self.t_feat.shape[1], self.feat_embed_dim)
        coo = _coo_override(config, "mm_adj", self.device)
        if coo is None:
            # freedom.py:64-77: w * image_adj + (1-w) * text_adj; duplicates are summed by SpMM
            ri, ci, vi = freedom_knn_coo(self.v_feat, self.knn_k, _knn_override(config, "image"))
            rt, ct, vt = freedom_knn_coo(self.t_feat, self.knn_k, _knn_override(config, "text"))
            w = self.mm_image_weight
            coo = (torch.cat([ri, rt]), torch.cat([ci, ct]), torch.cat([w * vi, (1.0 - w) * vt]))
        self.mm_adj = G.csr_from_coo(*coo, self.n_items, self.n_items)

    def pre_epoch_processing(self):
        """freedom.py:130-145."""
        if self.dropout <= .0:
            self.masked_adj = self.norm_adj
            return
        degree_len = int(self.edge_values.size(0) * (1. - self.dropout))
        # the reference draws on whatever device its model lives on (freedom.py:46, 136): the CUDA
        # generator on a GPU. edge_dropout_rng = "cpu" draws from torch's CPU generator instead --
        # the stream a CPU run of the reference consumes, which is what its fixtures were made with
        ev = self.edge_values.cpu() if self.edge_dropout_rng == "cpu" else self.edge_values
        degree_idx = torch.multinomial(ev, degree_len)
        self._set_masked_adj(self._masked_graph(degree_idx))

    def _masked_graph(self, keep_idx):
        keep_idx = keep_idx.to(self.device)
        return G.build_ui_graph(self._edge_u[keep_idx], self._edge_i[keep_idx], self.n_users,
                                self.n_items, "edge_f32")

    def forward(self, adj):
        h = self.item_id_embedding.weight
        for _ in range(self.n_layers):
            h = ops.spmm(self.mm_adj, h)
        ego = torch.cat((self.user_embedding.weight, self.item_id_embedding.weight), dim=0)
        out = ops.propagate_mean(adj, ego, self.n_ui_layers)
        u_g, i_g = _split(out, self.n_users)
        return u_g, i_g + h

    def _eval_forward(self):
        return self.forward(self.norm_adj)

    def calculate_loss(self, interaction):
        """freedom.py:191-212."""
        users, pos, neg = interaction[0], interaction[1], interaction[2]
        ua, ia = self.forward(self.masked_adj)
        B = users.shape[0]
        loss = ops.bpr(ua, ia, users, pos, neg)[0] / B
        mf_v = mf_t = 0.0
        if self.t_feat is not None:
            text_feats = self._project(self.text_embedding, self.text_trs)
            mf_t = ops.bpr(ua, text_feats, users, pos, neg)[0] / B
        if self.v_feat is not None:
            image_feats = self._project(self.image_embedding, self.image_trs)
            mf_v = ops.bpr(ua, image_feats, users, pos, neg)[0] / B
        return loss + self.reg_weight * (mf_t + mf_v)


# ============================================================================== MGCN / SMORE
class _MultiViewBase(GeneralRecommender):
    """Shared pieces of MGCN (mgcn.py) and SMORE (smore.py)."""

    def _init_ui(self):
        self.norm_adj = G.build_ui_graph(self._edge_u, self._edge_i, self.n_users, self.n_items,
                                         "f32")
        self.R, self.R_t = G.ui_blocks(self.norm_adj)

    def _item_graph(self, config, name, feat, k):
        coo = _coo_override(config, name, self.device)
        if coo is None:
            coo = knn_sym_coo(feat, k, _knn_override(config, name.split("_")[0]))
        return coo, G.csr_from_coo(*coo, self.n_items, self.n_items)

    # ---- independent branches of the forward on side streams ---------------------------------
    # The user-item propagation (4 latency-bound SpMMs), the text projection and the image
    # projection do not depend on one another until the side network; on one stream they run back
    # to back while each leaves most of the GPU idle. Forked onto side streams they overlap -- in
    # eager mode and, because the fork/join is captured, inside the CUDA graph of the training
    # step; autograd replays every backward node on the stream of its forward, so the backward
    # overlaps the same way.
    def _fork(self, idx):
        if not (getattr(self, "overlap_streams", False) and torch.device(self.device).type == "cuda"):
            return None
        streams = self.__dict__.setdefault("_aux_streams", {})
        if idx not in streams:
            streams[idx] = torch.cuda.Stream(device=self.device)
            # parameters used on a side stream get their AccumulateGrad on it: intended here
            hush = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
            if hush is not None:
                hush(False)
        s = streams[idx]
        s.wait_stream(torch.cuda.current_stream())
        return s

    @staticmethod
    def _join(s, *tensors):
        if s is None:
            return
        main = torch.cuda.current_stream()
        main.wait_stream(s)
        for t in tensors:
            t.record_stream(main)

    def _view(self, x, item_graph):
        for _ in range(self.n_layers):
            x = ops.spmm(item_graph, x)
        return torch.cat([ops.spmm(self.R, x), x], dim=0)

    def _views(self, xs, item_graphs):
        """`_view` for all modalities at once: the item-item hops of the views are independent of
        one another (and so are their user-side R products), so each hop is ONE launch."""
        xs = list(xs)
        if self.n_layers >= 1 and 1 < len(xs) <= 4 and self.R.t is not None:
            return ops.modality_views(item_graphs, self.R, self.n_layers, xs)
        for _ in range(self.n_layers):
            xs = ops.spmm_multi(item_graphs, xs)
        us = ops.spmm_multi([self.R] * len(xs), xs)
        return [torch.cat([u, x], dim=0) for u, x in zip(us, xs)]

    def _batch_arange(self, B, device):
        """arange(2 B) on the device, kept per batch size (index tensors of the compact gathers)."""
        cache = self.__dict__.setdefault("_arange_cache", {})
        key = (B, str(device))
        if key not in cache:
            cache[key] = torch.arange(2 * B, dtype=torch.int64, device=device)
        return cache[key]

    def _batch_loss(self, all_e, side, content, B, temperature):
        """BPR + InfoNCE of mgcn.py:233-253 / smore.py:389-411 on compact tables [3 B, d] whose rows 0..B-1 are
        the batch users, B..2B-1 the positive and 2B..3B-1 the negative items: the "user table" of the gathers
        has B rows, the "item table" 2 B."""
        ar = self._batch_arange(B, all_e.device)
        o = ops.bpr_table(all_e, B, ar[:B], ar[:B], ar[B:])
        cl = ops.infonce_pair(side, content, B, ar[:B], ar[:B], temperature, reduce=False)
        return ops.loss_head(o, cl, B, self.reg_weight, self.batch_size, self.cl_loss)

    def _eval_forward(self):
        return self.forward(self.norm_adj)


class MGCN(_MultiViewBase):
    """models/mgcn.py:21-263."""

    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.sparse = True
        self.cl_loss = config["cl_loss"]
        self.n_ui_layers = config["n_ui_layers"]
        self.embedding_dim = config["embedding_size"]
        self.knn_k = config["knn_k"]
        self.n_layers = config["n_layers"]
        self.reg_weight = config["reg_weight"]
        d = self.embedding_dim
        self.user_embedding = nn.Embedding(self.n_users, d)
        self.item_id_embedding = nn.Embedding(self.n_items, d)
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.item_id_embedding.weight)
        self._init_ui()
        self.image_embedding = self._feature_table(self.v_feat)
        _, self.image_original_adj = self._item_graph(config, "image_adj", self.v_feat, self.knn_k)
        self.text_embedding = self._feature_table(self.t_feat)
        _, self.text_original_adj = self._item_graph(config, "text_adj", self.t_feat, self.knn_k)
        self.image_trs = ops.Linear(self.v_feat.shape[1], d)
        self.text_trs = ops.Linear(self.t_feat.shape[1], d)
        self.softmax = nn.Softmax(dim=-1)
        # query_common.2 is the Linear(d, 1) of the attention head: a plain nn.Linear holder, its
        # weight is consumed as the row-dot vector of ops.mgcn_fuse
        self.query_common = ops.DenseStack(ops.Linear(d, d), nn.Tanh(), nn.Linear(d, 1, bias=False))
        self.gate_v = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.gate_t = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.gate_image_prefer = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.gate_text_prefer = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.tau = 0.5
        # training evaluates the row-local tail (attention fuser, preference gates, + content) on the rows of the
        # batch only, the user rows R x' of the views for the batch users only -- see SMORE.batch_rows
        self.batch_rows = bool(config.get("batch_rows", os.environ.get("MMREC_BATCH_ROWS", "1") != "0"))

    def forward(self, adj, train=False):
        """mgcn.py:146-208."""
        all_e, side, content = self._forward_full(adj)
        u, i = _split(all_e, self.n_users)
        return (u, i, side, content) if train else (u, i)

    def _forward_full(self, adj, batch=None):
        """(all_embeds, side_embeds, content_embeds) over all nodes; with `batch` = (users, pos, neg) over the
        3 B rows users | n_users + pos | n_users + neg of the batch."""
        image_feats = self._project(self.image_embedding, self.image_trs)
        text_feats = self._project(self.text_embedding, self.text_trs)
        item = self.item_id_embedding.weight
        gv, gt = ops.dense_stack_batch((self.gate_v, self.gate_t), (image_feats, text_feats))
        image_item = item * gv
        text_item = item * gt
        ego = torch.cat([self.user_embedding.weight, item], dim=0)
        content = ops.propagate_mean(adj, ego, self.n_ui_layers)
        item_graphs = (self.image_original_adj, self.text_original_adj)
        if batch is not None:
            xs = [image_item, text_item]
            for _ in range(self.n_layers):
                xs = ops.spmm_multi(item_graphs, xs)
            _, content, (image_embeds, text_embeds) = ops.gather_batch_views(self.R, xs, content, *batch)
        else:
            image_embeds, text_embeds = self._views((image_item, text_item), item_graphs)
        # attention fuser (mgcn.py:188-205): the two tanh layers and the two preference gates as one
        # batched launch each, everything after them (Linear(d, 1), softmax, common / specific
        # split, / 3, + content) in one kernel
        q0 = self.query_common[0]
        hi, ht = ops.dense_act_batch((image_embeds, text_embeds), (q0, q0), "tanh")
        pi, pt = ops.dense_act_batch((content, content), (self.gate_image_prefer[0], self.gate_text_prefer[0]),
                                     "sigmoid")
        all_e, side = ops.mgcn_fuse(hi, ht, self.query_common[2].weight, image_embeds, text_embeds, pi, pt, content)
        return all_e, side, content

    def calculate_loss(self, interaction):
        """mgcn.py:233-253."""
        users, pos, neg = interaction[0], interaction[1], interaction[2]
        if self.batch_rows and users.is_cuda:
            all_e, side, content = self._forward_full(self.norm_adj, batch=(users, pos, neg))
            return self._batch_loss(all_e, side, content, int(users.shape[0]), 0.2)
        all_e, side, content = self._forward_full(self.norm_adj)
        o = ops.bpr_table(all_e, self.n_users, users, pos, neg)
        cl = ops.infonce_pair(side, content, self.n_users, users, pos, 0.2, reduce=False)
        return ops.loss_head(o, cl, users.shape[0], self.reg_weight, self.batch_size, self.cl_loss)


class SMORE(_MultiViewBase):
    """models/smore.py:24-449 (this fork: residual injection, unit-magnitude spectral weights,
    mirror-gradient flags read by the trainer)."""

    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.sparse = True
        self.cl_loss = config["cl_loss"]
        self.n_ui_layers = config["n_ui_layers"]
        self.embedding_dim = config["embedding_size"]
        self.n_layers = config["n_layers"]
        self.reg_weight = config["reg_weight"]
        self.image_knn_k = config["image_knn_k"]
        self.text_knn_k = config["text_knn_k"]
        self.dropout_rate = config["dropout_rate"]
        self.dropout = nn.Dropout(p=self.dropout_rate)
        d = self.embedding_dim
        self.user_embedding = nn.Embedding(self.n_users, d)
        self.item_id_embedding = nn.Embedding(self.n_items, d)
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.item_id_embedding.weight)
        self._init_ui()
        self.image_embedding = self._feature_table(self.v_feat)
        img_coo, self.image_original_adj = self._item_graph(config, "image_adj", self.v_feat,
                                                            self.image_knn_k)
        self.text_embedding = self._feature_table(self.t_feat)
        txt_coo, self.text_original_adj = self._item_graph(config, "text_adj", self.t_feat,
                                                           self.text_knn_k)
        fus = _coo_override(config, "fusion_adj", self.device)
        if fus is None:
            fus = max_pool_fusion_coo(img_coo, txt_coo, self.n_items)
        self.fusion_adj = G.csr_from_coo(*fus, self.n_items, self.n_items)
        self.image_trs = ops.Linear(self.v_feat.shape[1], d)
        self.text_trs = ops.Linear(self.t_feat.shape[1], d)
        self.softmax = nn.Softmax(dim=-1)
        self.query_v = ops.DenseStack(ops.Linear(d, d), nn.Tanh(), ops.Linear(d, d, bias=False))
        self.query_t = ops.DenseStack(ops.Linear(d, d), nn.Tanh(), ops.Linear(d, d, bias=False))
        self.gate_v = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.gate_t = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.gate_f = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.gate_image_prefer = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.gate_text_prefer = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.gate_fusion_prefer = ops.DenseStack(ops.Linear(d, d), nn.Sigmoid())
        self.image_complex_weight = nn.Parameter(torch.randn(1, d // 2 + 1, 2, dtype=torch.float32))
        self.text_complex_weight = nn.Parameter(torch.randn(1, d // 2 + 1, 2, dtype=torch.float32))
        self.fusion_complex_weight = nn.Parameter(torch.randn(1, d // 2 + 1, 2, dtype=torch.float32))
        # smore.py:128-146
        self.mg_enable = bool(config.get("mg_enable", True))
        self.mg_interval = int(config.get("mg_interval", 3))
        self.mg_alpha = float(config.get("mg_alpha", 0.5))
        self.mg_beta = float(config.get("mg_beta", 0.2))
        self.mg_verbose = bool(config.get("mg_verbose", False))
        self.global_step = 0
        self.inject_mode = config.get("inject_mode", "residual")
        self.inject_scale = float(config.get("inject_scale", 0.7))
        self.spectral_weight_norm = bool(config.get("spectral_weight_norm", True))
        self.cl_temp = float(config.get("cl_temp", 0.2))
        self.overlap_streams = bool(config.get("overlap_streams", os.environ.get("MMREC_OVERLAP", "1") != "0"))
        # nn.Dropout of smore.py:331-333 generated inside the preference-module kernels (no mask
        # tensors): stream key = (seed, forward-call index, device counter). The trainer points
        # `dropout_counter` at FusedAdam's device-side update count, so replays of a captured step
        # draw fresh masks; without it (eager use) the call index alone advances the stream.
        self.fused_dropout = bool(config.get("fused_dropout", os.environ.get("MMREC_FUSED_DROPOUT", "1") != "0"))
        # Training evaluates the (row-local) preference module on the rows of the batch only: smore.py:395-407
        # consumes ua[users], ia[pos], ia[neg], side / content [users], [pos] and nothing else of it, and a row
        # nothing consumes gets a zero gradient (ops.gather_batch_rows, csrc/batch_rows.cu). Same loss, same
        # gradients; MMREC_BATCH_ROWS=0 / config["batch_rows"] = False evaluates all rows like the reference.
        self.batch_rows = bool(config.get("batch_rows", os.environ.get("MMREC_BATCH_ROWS", "1") != "0"))
        # Opt-in (config["batch_views"] / MMREC_BATCH_VIEWS=1): also form the user rows R x' of the modality views
        # for the users of the batch only (ops.gather_batch_views) instead of one more SpMM over all users and R^T
        # in the backward. Same loss and gradients (tested); measured slower at Baby / Sports size (2.18 vs 2.10 ms,
        # 3.65 vs 3.58 ms per step: a batch draws users in proportion to their interactions, so its 2 048 user rows
        # hold about as many non-zeros as a third of R and walk them with less parallelism) and 1 % faster at
        # Clothing d = 128 (8.53 vs 8.63 ms) -- off by default. MGCN, whose tail is lighter, gains 14 % from it.
        self.batch_views = self.batch_rows and bool(config.get("batch_views", os.environ.get("MMREC_BATCH_VIEWS", "0") != "0"))
        self.dropout_counter = None
        self._drop_seed = int(config.get("seed", 999))
        self._drop_calls = 0

    def _dropout_spec(self, row_ids=None, n_total=0):
        """(p, seed, counter[, row_ids, n_total]) of this forward call for ops.smore_side / ops.smore_combine."""
        self._drop_calls += 1
        if self.dropout_counter is None and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("SMORE: in-kernel dropout inside a captured step needs `dropout_counter` "
                               "(a device-side count that changes between replays; Trainer sets it)")
        spec = (self.dropout_rate, (self._drop_seed << 32) ^ (self._drop_calls * 0x9E3779B97F4A7C15), self.dropout_counter)
        return spec if row_ids is None else spec + (row_ids, int(n_total))

    def spectrum_convolution(self, image_embeds, text_embeds):
        """smore.py:209-252 without the band-energy .item() syncs (diagnostics only)."""
        # the [1, d/2+1, 2] parameters go in whole: indexing them here would add a select-backward
        # (zero fill + copy) per weight and pass to the autograd graph
        return ops.spectrum_convolution(image_embeds, text_embeds, self.image_complex_weight,
                                        self.text_complex_weight, self.fusion_complex_weight,
                                        self.spectral_weight_norm)

    def forward(self, adj, train=False):
        """smore.py:255-364."""
        all_e, side, content = self._forward_full(adj)
        u, i = _split(all_e, self.n_users)
        return (u, i, side, content) if train else (u, i)

    def _forward_full(self, adj, batch=None):
        """(all_embeds, side_embeds, content_embeds) over all nodes; with `batch` = (users, pos, neg) and
        `batch_rows` on, over the 3 B rows users | n_users + pos | n_users + neg of the batch instead."""
        import contextlib
        item = self.item_id_embedding.weight
        s_ui, s_txt = self._fork(0), self._fork(1)
        with torch.cuda.stream(s_ui) if s_ui is not None else contextlib.nullcontext():
            ego = torch.cat([self.user_embedding.weight, item], dim=0)
            content = ops.propagate_mean(adj, ego, self.n_ui_layers)
        with torch.cuda.stream(s_txt) if s_txt is not None else contextlib.nullcontext():
            text_feats = self._project(self.text_embedding, self.text_trs)
        image_feats = self._project(self.image_embedding, self.image_trs)
        self._join(s_txt, text_feats)
        image_conv, text_conv, fusion_conv = self.spectrum_convolution(image_feats, text_feats)
        if self.inject_mode == "mul":
            gv, gt, gf = ops.dense_stack_batch((self.gate_v, self.gate_t, self.gate_f),
                                               (image_conv, text_conv, fusion_conv))
            image_item, text_item, fusion_item = item * gv, item * gt, item * gf
        elif item.shape[1] % 4 == 0:
            gv, gt, gf = ops.dense_stack_batch((self.gate_v, self.gate_t, self.gate_f),
                                               (image_conv, text_conv, fusion_conv))
            image_item, text_item, fusion_item = ops.inject3(item, gv, gt, gf, self.inject_scale)
        else:
            image_item = item + self.inject_scale * self.gate_v(image_conv)
            text_item = item + self.inject_scale * self.gate_t(text_conv)
            fusion_item = item + self.inject_scale * self.gate_f(fusion_conv)
        item_graphs = (self.image_original_adj, self.text_original_adj, self.fusion_adj)
        if batch is not None and self.batch_views:
            # Everything from here on is row-local and only the batch rows are consumed: the item-item hops stay
            # dense, the user rows R x' of the three views are formed for the users of the batch only, together
            # with the gather of the item / content rows (gradients scatter back along the same non-zeros).
            xs = [image_item, text_item, fusion_item]
            for _ in range(self.n_layers):
                xs = ops.spmm_multi(item_graphs, xs)
            self._join(s_ui, content)
            n_total = int(content.shape[0])
            row_ids, content, (image_embeds, text_embeds, fusion_embeds) = ops.gather_batch_views(
                self.R, xs, content, batch[0], batch[1], batch[2])
        else:
            image_embeds, text_embeds, fusion_embeds = self._views((image_item, text_item, fusion_item), item_graphs)
            self._join(s_ui, content)
            row_ids, n_total = None, int(content.shape[0])
            if batch is not None:
                # everything below is row-local: keep the rows of the batch only (gradients scatter back)
                row_ids, (fusion_embeds, image_embeds, text_embeds, content) = ops.gather_batch_rows(
                    (fusion_embeds, image_embeds, text_embeds, content), batch[0], batch[1], batch[2], self.n_users)
        # modality-aware preference module (smore.py:321-341): one fused kernel for d = 32 / 64
        if ops.smore_side_supported(self.embedding_dim):
            masks = drop = None
            if self.training and self.dropout_rate > 0:
                if self.fused_dropout:
                    drop = self._dropout_spec(row_ids, n_total)
                else:
                    # the three nn.Dropout masks (smore.py:331-333) drawn in one call
                    masks = torch.nn.functional.dropout(
                        torch.ones(3, *content.shape, dtype=content.dtype, device=content.device),
                        p=self.dropout_rate, training=True)
            layers = (self.query_v[0], self.query_v[2], self.query_t[0], self.query_t[2],
                      self.gate_image_prefer[0], self.gate_text_prefer[0], self.gate_fusion_prefer[0])
            all_e, side = ops.smore_side(fusion_embeds, image_embeds, text_embeds, content, layers, masks, drop)
            return all_e, side, content
        if ops.smore_combine_supported(self.embedding_dim):
            # wide embeddings (d = 128): the seven Linear layers as tensor-core launches, everything
            # after them (two softmaxes, dropout, products, mean of three, + content) in one kernel
            zv, zt = self.query_v(fusion_embeds), self.query_t(fusion_embeds)
            gi, gt, gf = ops.dense_stack_batch(
                (self.gate_image_prefer, self.gate_text_prefer, self.gate_fusion_prefer), (content, content, content))
            masks = drop = None
            if self.training and self.dropout_rate > 0:
                if self.fused_dropout:
                    drop = self._dropout_spec(row_ids, n_total)
                else:
                    masks = torch.nn.functional.dropout(
                        torch.ones(3, *content.shape, dtype=content.dtype, device=content.device),
                        p=self.dropout_rate, training=True)
            all_e, side = ops.smore_combine(zv, zt, image_embeds, text_embeds, fusion_embeds, content, gi, gt, gf,
                                            masks, drop)
            return all_e, side, content
        agg_image = self.softmax(self.query_v(fusion_embeds)) * image_embeds
        agg_text = self.softmax(self.query_t(fusion_embeds)) * text_embeds
        pi, pt, pf = ops.dense_stack_batch(
            (self.gate_image_prefer, self.gate_text_prefer, self.gate_fusion_prefer), (content, content, content))
        image_prefer, text_prefer, fusion_prefer = self.dropout(pi), self.dropout(pt), self.dropout(pf)
        side = torch.mean(torch.stack([image_prefer * agg_image, text_prefer * agg_text,
                                       fusion_prefer * fusion_embeds]), dim=0)
        return content + side, side, content

    def calculate_loss(self, interaction):
        """smore.py:389-411 (global_step is bumped here, which is what makes every step after
        the second a mirror-gradient step in the trainer -- SURVEY 3.2)."""
        users, pos, neg = interaction[0], interaction[1], interaction[2]
        if self.batch_rows and users.is_cuda:
            # compact tables [3 B, d]: rows 0..B-1 = users, B..2B-1 = positive items, 2B..3B-1 = negative items --
            # the "user table" of the gathers below has B rows, the "item table" 2 B
            all_e, side, content = self._forward_full(self.norm_adj, batch=(users, pos, neg))
            self.global_step += 1
            return self._batch_loss(all_e, side, content, int(users.shape[0]), self.cl_temp)
        all_e, side, content = self._forward_full(self.norm_adj)
        self.global_step += 1
        o = ops.bpr_table(all_e, self.n_users, users, pos, neg)
        cl = ops.infonce_pair(side, content, self.n_users, users, pos, self.cl_temp, reduce=False)
        return ops.loss_head(o, cl, users.shape[0], self.reg_weight, self.batch_size, self.cl_loss)


MODELS = {"LightGCN": LightGCN, "LayerGCN": LayerGCN, "FREEDOM": FREEDOM, "MGCN": MGCN,
          "SMORE": SMORE}


def get_model(model_name):
    """utils/utils.py:28-41."""
    return MODELS[model_name]
