"""Shared by the GPU parity tests and __graft_entry__.smoke(): build our model on the tiny
synthetic dataset, load the reference's initial parameters from the golden fixture and compare
forward / loss / gradients / fused top-K against the oracle and the golden vectors."""
import numpy as np
import torch

from conftest import golden, pkg, TINY
from oracle import build as obuild
from oracle import models as omodels
from oracle import ops as oops

GOLDEN_OVERRIDES = {"train_batch_size": 512, "eval_batch_size": 64, "epochs": 2}
MODEL_OVERRIDES = {
    "LightGCN": {}, "LayerGCN": {"dropout": 0.0, "reg_weight": 1e-2},
    "FREEDOM": {"dropout": 0.8, "reg_weight": 1e-3}, "MGCN": {"cl_loss": 0.01},
    "SMORE": {"dropout_rate": 0.0},
}
TAGS = {"LightGCN": "tiny_lightgcn", "LayerGCN": "tiny_layergcn", "FREEDOM": "tiny_freedom",
        "MGCN": "tiny_mgcn", "SMORE": "tiny_smore"}
GRAPH_KEYS = {"FREEDOM": {"mm_adj": "mm_adj"},
              "MGCN": {"image_adj": "image_original_adj", "text_adj": "text_original_adj"},
              "SMORE": {"image_adj": "image_original_adj", "text_adj": "text_original_adj",
                        "fusion_adj": "fusion_adj"}}


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-6))


def make_env(model_name, device, tag=None, overrides=None, inject_graphs=True):
    synth, cfgm, data_m, models = pkg("synth"), pkg("config"), pkg("data"), pkg("models")
    g = golden(tag or TAGS[model_name])
    data = synth.make_dataset("tiny", **TINY)
    cd = dict(GOLDEN_OVERRIDES)
    cd.update(MODEL_OVERRIDES[model_name])
    cd.update(overrides or {})
    cd.update({"device": torch.device(device), "data_path": None})
    if model_name not in ("LightGCN", "LayerGCN"):
        cd.update({"v_feat": data.image_feat, "t_feat": data.text_feat})
        if inject_graphs:
            cd["item_graphs"] = {ours: (g[f"adj/{theirs}/idx"][0], g[f"adj/{theirs}/idx"][1],
                                        g[f"adj/{theirs}/val"])
                                 for ours, theirs in GRAPH_KEYS[model_name].items()}
    else:
        cd["is_multimodal_model"] = False
    config = cfgm.Config(model_name, "tiny", cd)
    ds = data_m.RecDataset(config, data.users, data.items, data.labels)
    tr, va, te = ds.split()
    train_data = data_m.TrainDataLoader(config, tr, batch_size=config["train_batch_size"], shuffle=True)
    valid_data = data_m.EvalDataLoader(config, va, additional_dataset=tr, batch_size=config["eval_batch_size"])
    test_data = data_m.EvalDataLoader(config, te, additional_dataset=tr, batch_size=config["eval_batch_size"])
    cfgm.init_seed(config["seed"])
    train_data.pretrain_setup()
    model = models.get_model(model_name)(config, train_data).to(config["device"])
    return dict(config=config, model=model, train=train_data, valid=valid_data, test=test_data,
                golden=g, data=data, tr=tr)


def golden_params(g):
    return {k[len("param0/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param0/")}


def oracle_cfg(config):
    keys = ["n_layers", "reg_weight", "n_mm_layers", "n_ui_layers", "knn_k", "mm_image_weight",
            "cl_loss", "train_batch_size", "image_knn_k", "text_knn_k"]
    return {k: config[k] for k in keys if config[k] is not None}


def run_model_parity(model_name, device="cuda:0", tol=1e-5, tag=None, overrides=None):
    env = make_env(model_name, device, tag=tag, overrides=overrides)
    model, g, config, data = env["model"], env["golden"], env["config"], env["data"]
    rep = {}
    # 1. same seed -> same initial parameters as the reference (RNG consumption order)
    P0 = golden_params(g)
    sd = model.state_dict()
    rep["init_bit_exact"] = all(torch.equal(sd[k].cpu(), v) for k, v in P0.items())
    model.load_state_dict({k: v.to(device) for k, v in P0.items()}, strict=True)
    # 2. oracle graphs (CPU) with the reference's item graphs
    tu, ti = env["tr"].users, env["tr"].items
    img = data.image_feat if model_name in GRAPH_KEYS else None
    ocfg = oracle_cfg(config)
    G, parts = obuild.build_graphs(model_name, tu, ti, data.n_users, data.n_items, img, data.text_feat, ocfg)
    adj = getattr(model, "norm_adj_matrix", None) or model.norm_adj
    r, c, v = adj.to_torch_coo()
    orr, occ, ov = parts["norm_adj"]
    rep["adj_bit_exact"] = bool(np.array_equal(r, orr) and np.array_equal(c, occ) and
                                np.array_equal(v.view(np.uint32), ov.view(np.uint32)))
    # 3. eval forward vs oracle and vs the reference's own output
    model.eval()
    ue, ie = model.restore_embeddings()
    Pcpu = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
    with torch.no_grad():
        oue, oie = omodels.FORWARD[model_name](Pcpu, G, ocfg)
    rep["fwd_vs_oracle"] = max(rel_err(ue.cpu(), oue), rel_err(ie.cpu(), oie))
    rep["fwd_vs_reference"] = max(rel_err(ue.cpu(), g["eval_user_emb"]), rel_err(ie.cpu(), g["eval_item_emb"]))
    # 4. loss + gradients on the reference's first batch
    model.train()
    if model_name in ("LayerGCN", "FREEDOM") and config["dropout"] > 0:
        midx, mval = g["adj/masked_adj/idx"], g["adj/masked_adj/val"]
        n = data.n_users + data.n_items
        from oracle import graph as ograph
        G["masked_adj"] = ograph.to_torch_csr(midx[0], midx[1], mval, (n, n))
        keep = midx.shape[1] // 2
        eu, ei = model._edge_u.cpu().numpy(), model._edge_i.cpu().numpy()
        key = eu * data.n_items + ei
        kept = midx[0, :keep] * data.n_items + (midx[1, :keep] - data.n_users)
        keep_idx = torch.from_numpy(np.searchsorted(key, kept))
        model.masked_adj = model._masked_graph(keep_idx)
        r, c, v = model.masked_adj.to_torch_coo()
        o = np.lexsort((midx[1], midx[0]))
        rep["masked_adj_bit_exact"] = bool(
            np.array_equal(r, midx[0][o]) and np.array_equal(c, midx[1][o]) and
            np.array_equal(v.view(np.uint32), mval[o].view(np.uint32)))
    else:
        model.pre_epoch_processing()
        if "masked_adj" not in G and "norm_adj" in G:
            G["masked_adj"] = G["norm_adj"]
    batch = torch.from_numpy(g["batch0"]).to(device)
    model.zero_grad()
    loss = model.calculate_loss(batch)
    loss.backward()
    oloss = omodels.LOSS[model_name](Pcpu, G, ocfg, torch.from_numpy(g["batch0"]))
    oloss.backward()
    rep["loss_vs_oracle"] = abs(loss.item() - oloss.item()) / abs(oloss.item())
    rep["loss_vs_reference"] = abs(loss.item() - float(g["loss0"])) / abs(float(g["loss0"]))
    gerr_o = gerr_r = 0.0
    worst = None
    for name, p in model.named_parameters():
        if Pcpu[name].grad is None:
            continue
        assert p.grad is not None, name
        eo = rel_err(p.grad.cpu(), Pcpu[name].grad)
        if eo > gerr_o:
            gerr_o, worst = eo, name
        if "grad0/" + name in g.files:
            gerr_r = max(gerr_r, rel_err(p.grad.cpu(), g["grad0/" + name]))
    rep["grad_vs_oracle"], rep["grad_vs_reference"], rep["worst_grad"] = gerr_o, gerr_r, worst
    # 5. fused score + mask + top-K on the reference's eval batch
    model.eval()
    users = torch.from_numpy(g["eval_batch_users"]).to(device)
    mask = g["eval_batch_mask"]
    rowptr = np.concatenate(([0], np.cumsum(np.bincount(mask[0], minlength=len(users))))).astype(np.int32)
    o = np.lexsort((mask[1], mask[0]))
    cols = mask[1][o].astype(np.int32)
    k = max(config["topk"])
    ids, vals = pkg("ops").score_mask_topk(ue, users, ie, k, torch.from_numpy(rowptr).to(device),
                                           torch.from_numpy(cols).to(device), return_scores=True)
    oscores = oops.full_sort_scores(ue.cpu(), ie.cpu(), users.cpu())     # oracle on OUR embeddings
    want = oops.mask_topk(oscores, torch.from_numpy(mask[0]), torch.from_numpy(mask[1]), k).numpy()
    rep["topk_id_mismatch"] = int((ids.cpu().numpy() != want).sum())
    ref = g["eval_topk_ref"]
    rep["topk_vs_reference_mismatch"] = int((ids.cpu().numpy() != ref).sum())
    rep["ok"] = bool(rep["adj_bit_exact"] and rep.get("masked_adj_bit_exact", True) and
                     rep["fwd_vs_oracle"] < tol and rep["fwd_vs_reference"] < 2 * tol and
                     rep["loss_vs_oracle"] < tol and rep["loss_vs_reference"] < tol and
                     rep["grad_vs_oracle"] < 5 * tol and rep["topk_id_mismatch"] == 0)
    return rep
