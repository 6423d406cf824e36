"""A few eager (non-graph) steady-state SMORE training steps + one evaluation pass: the command
ncu wraps. Usage: python scripts/ncu_step.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
env = bench.build_env("cuda:0", overrides={"cuda_graph": False})
trainer = bench.pkg("trainer").Trainer(env["config"], env["model"])
batches = bench.take_batches(env["train"], 3 + steps)
env["model"].train()
for b in batches:
    trainer._train_batch(b)
torch.cuda.synchronize()
trainer.evaluate_topk(env["valid"])
torch.cuda.synchronize()
print("done")
