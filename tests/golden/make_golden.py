"""Generate golden vectors by executing the UNMODIFIED reference (/root/reference/src) in-process.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference ships no tests or fixtures, so parity is pinned by running its
own model / loader / trainer / evaluator classes on the deterministic synthetic "tiny" dataset
(`synth.make_dataset('tiny', image_dim=128, text_dim=48)`) with seed 999 on CPU. Shims (SURVEY.md
section 8c): a stub `matplotlib`, `torch_scatter.scatter_add` -> index_add_, `Tensor.cuda` no-op.
Nothing from the reference is copied; only the tensors it produces are stored.
"""
import importlib
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
SCRATCH = "/tmp/mmrec_golden"

sys.path.insert(0, REPO)
pkg = importlib.import_module("recommendar-systems_b200")
synth = importlib.import_module("recommendar-systems_b200.synth")

TINY = dict(image_dim=128, text_dim=48)
BASE_OVERRIDES = {"use_gpu": False, "train_batch_size": 512, "eval_batch_size": 64, "epochs": 2,
                  "data_path": os.path.join(SCRATCH, "data") + "/", "mg_verbose": False,
                  "diag_gate": False, "diag_spectrum": False, "diag_grad": False}


def install_shims():
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    ts = types.ModuleType("torch_scatter")

    def scatter_add(src, index, dim=0, dim_size=None):
        out = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        return out.index_add_(0, index, src)
    ts.scatter_add = scatter_add
    sys.modules.setdefault("torch_scatter", ts)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self


def prepare_scratch():
    os.makedirs(os.path.join(SCRATCH, "configs", "dataset"), exist_ok=True)
    for name in ("overall.yaml", "model", "mg.yaml"):
        dst = os.path.join(SCRATCH, "configs", name)
        if not os.path.lexists(dst):
            os.symlink(os.path.join(REF, "configs", name), dst)
    with open(os.path.join(SCRATCH, "configs", "dataset", "tiny.yaml"), "w") as f:
        f.write("USER_ID_FIELD: userID\nITEM_ID_FIELD: itemID\nTIME_FIELD: timestamp\n"
                "filter_out_cod_start_users: True\ninter_file_name: 'tiny.inter'\n"
                "vision_feature_file: 'image_feat.npy'\ntext_feature_file: 'text_feat.npy'\n"
                "field_separator: \"\\t\"\n")
    data = synth.make_dataset("tiny", **TINY)
    d = synth.write_reference_layout(data, os.path.join(SCRATCH, "data"))
    for fn in os.listdir(d):            # stale kNN caches silently win (smore.py:56-72)
        if fn.endswith(".pt"):
            os.remove(os.path.join(d, fn))
    os.chdir(SCRATCH)
    sys.path.insert(0, REF)
    return data


def sp_parts(t):
    return t._indices().cpu().numpy().copy(), t._values().cpu().numpy().copy()


def build(model_name, overrides):
    from utils.configurator import Config
    from utils.dataset import RecDataset
    from utils.dataloader import TrainDataLoader, EvalDataLoader
    from utils.utils import init_seed, get_model, get_trainer
    cd = dict(BASE_OVERRIDES)
    cd.update(overrides)
    config = Config(model_name, "tiny", cd)
    dataset = RecDataset(config)
    str(dataset)
    tr, va, te = dataset.split()
    str(tr), str(va), str(te)
    train_data = TrainDataLoader(config, tr, batch_size=config["train_batch_size"], shuffle=True)
    valid_data = EvalDataLoader(config, va, additional_dataset=tr,
                                batch_size=config["eval_batch_size"])
    test_data = EvalDataLoader(config, te, additional_dataset=tr,
                               batch_size=config["eval_batch_size"])
    for k in config["hyper_parameters"]:
        v = config[k]
        if isinstance(v, list):
            config[k] = cd.get(k, v[0]) if not isinstance(cd.get(k), list) else v[0]
    for k, v in overrides.items():
        config[k] = v
    init_seed(config["seed"])
    train_data.pretrain_setup()
    ddir = os.path.join(SCRATCH, "data", "tiny")
    for fn in os.listdir(ddir):
        if fn.endswith(".pt"):
            os.remove(os.path.join(ddir, fn))
    model = get_model(model_name)(config, train_data).to(config["device"])
    trainer = get_trainer()(config, model, False)
    return config, model, trainer, train_data, valid_data, test_data


ADJ_ATTRS = ["norm_adj_matrix", "norm_adj", "R", "mm_adj", "image_original_adj",
             "text_original_adj", "fusion_adj"]


def capture(model_name, overrides, tag):
    out = {}
    config, model, trainer, train_data, valid_data, test_data = build(model_name, overrides)
    out["all_items_shuffled"] = np.asarray(train_data.all_items, dtype=np.int64)
    for n, p in model.named_parameters():
        out["param0/" + n] = p.detach().numpy().copy()
    for a in ADJ_ATTRS:
        t = getattr(model, a, None)
        if t is not None and torch.is_tensor(t) and t.is_sparse:
            out[f"adj/{a}/idx"], out[f"adj/{a}/val"] = sp_parts(t)
    if hasattr(model, "edge_values"):
        out["edge_values"] = model.edge_values.numpy().copy()
        out["edge_indices"] = model.edge_indices.numpy().copy()
    # ---- one epoch prologue: pre_epoch_processing (edge dropout) then two batches
    model.pre_epoch_processing()
    if getattr(model, "masked_adj", None) is not None:
        out["adj/masked_adj/idx"], out["adj/masked_adj/val"] = sp_parts(model.masked_adj)
    it = iter(train_data)
    batches = [next(it).clone(), next(it).clone()]
    out["batch0"], out["batch1"] = batches[0].numpy(), batches[1].numpy()
    train_data.pr = 0
    # ---- loss + grads on batch0 (train mode)
    model.train()
    model.zero_grad()
    if hasattr(model, "global_step"):
        gs = model.global_step
    torch_state = torch.get_rng_state()
    loss = model.calculate_loss(batches[0])
    loss.backward()
    out["loss0"] = np.asarray(loss.item(), dtype=np.float64)
    for n, p in model.named_parameters():
        if p.grad is not None:
            out["grad0/" + n] = p.grad.numpy().copy()
    torch.set_rng_state(torch_state)
    if hasattr(model, "global_step"):
        model.global_step = gs
    model.zero_grad()
    # ---- train-mode forward internals where they exist
    with torch.no_grad():
        model.eval()
        if model_name in ("SMORE", "MGCN", "FREEDOM"):
            adj = model.norm_adj
            ue, ie = model.forward(adj)
        else:
            model.forward_adj = model.norm_adj_matrix
            ue, ie = model.forward()
        out["eval_user_emb"], out["eval_item_emb"] = ue.numpy().copy(), ie.numpy().copy()
        if model_name == "SMORE":
            img = model.image_trs(model.image_embedding.weight)
            txt = model.text_trs(model.text_embedding.weight)
            ic, tc, fc = model.spectrum_convolution(img, txt)
            out["spec/image_feats"], out["spec/text_feats"] = img.numpy().copy(), txt.numpy().copy()
            out["spec/image_conv"], out["spec/text_conv"], out["spec/fusion_conv"] = \
                ic.numpy().copy(), tc.numpy().copy(), fc.numpy().copy()
        # ---- eval batch 0: scores, masked top-K (reference torch.topk), metrics
        ev = iter(valid_data)
        b = next(ev)
        valid_data.pr = 0
        valid_data.inter_pr = 0
        scores = model.full_sort_predict(b)
        out["eval_batch_users"], out["eval_batch_mask"] = b[0].numpy().copy(), b[1].numpy().copy()
        out["eval_scores"] = scores.numpy().copy()
        masked = scores.clone()
        masked[b[1][0], b[1][1]] = -1e10
        out["eval_topk_ref"] = torch.topk(masked, max(config["topk"]), dim=-1)[1].numpy().copy()
    # ---- two epochs through the reference Trainer (fit()'s loop, trainer.py:408-481)
    init_rng = (random.getstate(), np.random.get_state(), torch.get_rng_state())
    losses, valids, tests, valid_raw = [], [], [], []
    for epoch in range(2):
        model.pre_epoch_processing()
        train_loss, _ = trainer._train_epoch(train_data, epoch)
        trainer.lr_scheduler.step()
        losses.append(float(train_loss))
        vr = trainer.evaluate(valid_data)
        tr_ = trainer.evaluate(test_data)
        valids.append([vr[k] for k in sorted(vr)])
        tests.append([tr_[k] for k in sorted(tr_)])
    out["fit/metric_keys"] = np.asarray(sorted(vr))
    out["fit/train_loss"] = np.asarray(losses)
    out["fit/valid"], out["fit/test"] = np.asarray(valids), np.asarray(tests)
    for n, p in model.named_parameters():
        if n.split(".")[0] in ("user_embeddings", "item_embeddings", "user_embedding",
                               "item_id_embedding", "image_trs", "image_complex_weight"):
            out["fit/param/" + n] = p.detach().numpy().copy()
    if hasattr(model, "global_step"):
        out["fit/global_step"] = np.asarray(model.global_step)
    # unrounded metrics for the final model on valid
    with torch.no_grad():
        model.eval()
        mats = []
        for b in valid_data:
            s = model.full_sort_predict(b)
            s[b[1][0], b[1][1]] = -1e10
            mats.append(torch.topk(s, max(config["topk"]), dim=-1)[1])
        topk_index = torch.cat(mats, 0).numpy()
        pos_items = valid_data.get_eval_items()
        hits = np.asarray([[i in set(m.tolist()) for i in n] for m, n in zip(pos_items, topk_index)])
        raw = trainer.evaluator._calculate_metrics(valid_data.get_eval_len_list(), hits)
        out["fit/valid_topk"] = topk_index
        out["fit/valid_metrics_raw"] = raw
        out["fit/metric_names"] = np.asarray(trainer.evaluator.metrics)
    path = os.path.join(HERE, f"{tag}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB, "
          f"loss0={out['loss0']:.6f} fit_loss={losses}")


def micro_vectors():
    """Known-answer micro-vectors on a 3-user x 2-item graph, produced by
    calling the reference functions unbound."""
    from models.layergcn import LayerGCN
    from models.mgcn import MGCN
    from models.freedom import FREEDOM
    from models.smore import SMORE
    from scipy.sparse import coo_matrix
    from utils import metrics as ref_metrics
    out = {}
    u = np.array([0, 0, 1, 2]); i = np.array([0, 1, 1, 0])
    inter = coo_matrix((np.ones(4), (u, i)), shape=(3, 2)).astype(np.float32)
    ns = types.SimpleNamespace(n_users=3, n_items=2, n_nodes=5, interaction_matrix=inter)
    adj = LayerGCN.get_norm_adj_mat(ns)
    out["layergcn_adj_idx"], out["layergcn_adj_val"] = sp_parts(adj)
    ns2 = types.SimpleNamespace(n_users=3, n_items=2, interaction_matrix=inter)
    csr = MGCN.get_adj_mat(ns2).tocoo()
    out["mgcn_adj_idx"] = np.vstack([csr.row, csr.col]).astype(np.int64)
    out["mgcn_adj_val"] = csr.data.astype(np.float32)
    R = ns2.R.tocoo()
    out["mgcn_R_idx"] = np.vstack([R.row, R.col]).astype(np.int64)
    out["mgcn_R_val"] = R.data.astype(np.float32)
    edges = torch.from_numpy(np.vstack([u, i])).long()
    out["edge_norm_val"] = LayerGCN._normalize_adj_m(None, edges, torch.Size((3, 2))).numpy()
    feats = torch.tensor([[1., 0, 0], [.9, .1, 0], [0, 1., 0]])
    ns3 = types.SimpleNamespace(knn_k=2, device=torch.device("cpu"))
    ns3.compute_normalized_laplacian = lambda ind, sz: FREEDOM.compute_normalized_laplacian(ns3, ind, sz)
    ind, kadj = FREEDOM.get_knn_adj_mat(ns3, feats)
    out["freedom_knn_idx"], out["freedom_knn_val"] = sp_parts(kadj)
    # LayerGCN.forward, n_layers=2
    x0 = torch.tensor([[1., 0], [0, 1], [1, 1], [1, 2], [2, 1]])
    ns4 = types.SimpleNamespace(n_users=3, n_items=2, n_layers=2, forward_adj=adj,
                                get_ego_embeddings=lambda: x0)
    ue, ie = LayerGCN.forward(ns4)
    out["layergcn_fwd_user"], out["layergcn_fwd_item"] = ue.numpy(), ie.numpy()
    users = torch.tensor([0, 2]); pos = torch.tensor([1, 0]); neg = torch.tensor([0, 1])
    out["layergcn_bpr_sum"] = LayerGCN.bpr_loss(None, ue, ie, users, pos, neg).numpy()
    # light-gcn style mean over layers 0..2 with the same adjacency
    e, layers = x0, [x0]
    for _ in range(2):
        e = torch.sparse.mm(adj, e)
        layers.append(e)
    out["lightgcn_mean"] = torch.stack(layers, 1).mean(1).numpy()
    # metrics
    hits = np.array([[1, 0, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 0]], dtype=bool)
    pos_len = np.array([2, 1, 7])
    for name in ("recall", "ndcg", "precision", "map"):
        out["metric_" + name] = ref_metrics.metrics_dict[name](hits, pos_len)
    # spectrum convolution, d=8
    x = torch.arange(1., 9.).view(1, 8)
    y = torch.tensor([[1., -1, 2, 0, .5, 0, 0, 3]])
    w = torch.tensor([[[1., 1], [0, 2], [3, 0], [-1, 1], [2, -2]]])
    ns5 = types.SimpleNamespace(image_complex_weight=w, text_complex_weight=w,
                                fusion_complex_weight=w, spectral_weight_norm=True)
    ic, tc, fc = SMORE.spectrum_convolution(ns5, x, y)
    out["spec_x"], out["spec_y"], out["spec_w"] = x.numpy(), y.numpy(), w.numpy()
    out["spec_image_conv"], out["spec_text_conv"], out["spec_fusion_conv"] = \
        ic.numpy(), tc.numpy(), fc.numpy()
    # torch.topk tie order and cosine_similarity semantics of the installed torch
    out["topk_ties_idx"] = torch.topk(torch.tensor([1., 3, 3, 2, 3, 0]), 3)[1].numpy()
    out["cos_tiny"] = torch.nn.functional.cosine_similarity(
        torch.tensor([[1e-9, 0.]]), torch.tensor([[1e-9, 0.]]), dim=-1).numpy()
    np.savez_compressed(os.path.join(HERE, "micro.npz"), **out)
    print("wrote micro.npz:", {k: v.shape for k, v in out.items()})


def main():
    install_shims()
    prepare_scratch()
    micro_vectors()
    capture("LightGCN", {}, "tiny_lightgcn")
    capture("LayerGCN", {"dropout": 0.0, "reg_weight": 1e-2}, "tiny_layergcn")
    capture("LayerGCN", {"dropout": 0.1, "reg_weight": 1e-3}, "tiny_layergcn_drop")
    capture("FREEDOM", {"dropout": 0.8, "reg_weight": 1e-3}, "tiny_freedom")
    capture("MGCN", {"cl_loss": 0.01}, "tiny_mgcn")
    capture("SMORE", {"dropout_rate": 0.0}, "tiny_smore")
    capture("SMORE", {"dropout_rate": 0.0, "mg_enable": False}, "tiny_smore_nomg")


if __name__ == "__main__":
    main()
