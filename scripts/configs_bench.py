"""The other BASELINE.json configs (parity-test cases, not bench lines): device time per training
step (CUDA-graph replay) and full-rank evaluation throughput for each model / dataset shape."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

CASES = [("LayerGCN", "baby", {"is_multimodal_model": False}),
         ("FREEDOM", "sports", {}), ("MGCN", "sports", {}),
         ("SMORE", "clothing", {"embedding_size": 128}), ("SMORE", "sports", {})]
EXTRA = [("SMORE", "baby", {})]          # the headline config: only on request (A/B runs of env switches)
only = sys.argv[1:] or None
for model_name, shape, over in CASES + EXTRA:
    if (only and model_name + ":" + shape not in only) or (not only and (model_name, shape, over) in EXTRA):
        continue
    torch.cuda.empty_cache()
    env = bench.build_env("cuda:0", model_name=model_name, shape=shape, overrides=over)
    trainer = bench.pkg("trainer").Trainer(env["config"], env["model"])
    K, W = 20, 8
    batches = bench.take_batches(env["train"], W + K)
    env["model"].train()
    env["model"].pre_epoch_processing()
    for b in batches[:W]:
        trainer._train_batch_graphed(b)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for b in batches[W:]:
        trainer._train_batch_graphed(b)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / K
    n_eval = int(env["valid"].eval_u.shape[0])
    for _ in range(3):                       # eager pass, graph capture, first replay
        trainer.evaluate(env["valid"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        metrics = trainer.evaluate(env["valid"])
    torch.cuda.synchronize()
    ev = (time.perf_counter() - t0) / 5
    B = env["config"]["train_batch_size"]
    n_train = len(env["tr"])
    print(json.dumps({"model": model_name, "shape": shape, "d": env["config"]["embedding_size"],
                      "ms_per_step": round(ms, 3), "interactions_per_s": round(B / ms * 1e3),
                      "train_epoch_s": round(-(-n_train // B) * ms / 1e3, 3), "eval_users": n_eval,
                      "eval_users_per_s": round(n_eval / ev), "recall@20": metrics.get("recall@20")}), flush=True)
    del env, trainer, batches
