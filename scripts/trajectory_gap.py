"""How close is the two-epoch trainer trajectory to the reference's fixture? Prints the actual deviations (losses,
unrounded metrics vs fit/valid_metrics_raw, top-K id agreement, parameters) so that the test tolerances can be set
from data."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import pkg
from parity_util import make_env, golden_params
DEV = "cuda:0"
for model, tag, over in [("LayerGCN", "tiny_layergcn", {}), ("MGCN", "tiny_mgcn", {}), ("SMORE", "tiny_smore", {}),
                         ("SMORE", "tiny_smore_nomg", {"mg_enable": False}), ("FREEDOM", "tiny_freedom", {}),
                         ("LayerGCN", "tiny_layergcn_drop", {"dropout": 0.1, "reg_weight": 1e-3})]:
    for graph in (True, False):
        env = make_env(model, DEV, tag=tag, overrides=dict(over, cuda_graph=graph))
        g, m, train, valid, test = env["golden"], env["model"], env["train"], env["valid"], env["test"]
        m.load_state_dict({k: v.to(DEV) for k, v in golden_params(g).items()})
        if "fit/train_loss" not in g.files:
            print({"model": tag, "note": "no fit/ keys in the fixture"}); break
        m.pre_epoch_processing()            # the fixture's prologue drew one edge-dropout sample before its fit loop
        it = iter(train); next(it), next(it); train.pr = 0
        tr = pkg("trainer").Trainer(env["config"], m)
        losses, raws, topks = [], [], []
        for epoch in range(2):
            m.pre_epoch_processing()
            loss, _ = tr._train_epoch(train, epoch)
            tr.lr_scheduler.step()
            losses.append(float(loss))
            v = tr.evaluate(valid)
            raws.append(getattr(tr, "last_metrics_raw", None))
            topks.append(getattr(tr, "last_topk", None))
            tr.evaluate(test)
        out = {"model": tag, "graph": graph}
        ids = torch.cat(tr.evaluate_topk(valid), dim=0)
        rowptr, items = valid.gt_csr()
        raw = tr.evaluator._metrics_from_sums(pkg("ops").topk_metric_sums(ids, rowptr, items), ids.shape[0], valid)
        if "fit/valid_metrics_raw" in g.files:
            names = [str(x).lower() for x in g["fit/metric_names"]]
            ours = np.stack([raw[tr.evaluator.metrics.index(n)] for n in names], axis=0)
            out["metrics_raw_maxabs"] = float(np.abs(ours - g["fit/valid_metrics_raw"]).max())
            out["topk_rows_equal"] = float((ids.cpu().numpy() == g["fit/valid_topk"]).all(axis=1).mean())
            out["topk_ids_equal"] = float((ids.cpu().numpy() == g["fit/valid_topk"]).mean())
        out["valid_rounded_maxabs"] = float(np.abs(np.asarray([[v[str(k)] for k in g["fit/metric_keys"]]]) - g["fit/valid"][-1]).max())
        out["loss_rel"] = float(np.max(np.abs(np.asarray(losses) - g["fit/train_loss"]) / np.abs(g["fit/train_loss"])))
        pmax = 0.0
        for k in g.files:
            if k.startswith("fit/param/"):
                ours = m.state_dict()[k[len("fit/param/"):]].cpu().numpy()
                pmax = max(pmax, float(np.abs(ours - g[k]).max() / np.abs(g[k]).max()))
        out["param_rel_max"] = pmax
        print(out, flush=True)
