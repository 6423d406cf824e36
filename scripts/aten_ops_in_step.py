"""Which ATen ops (and shapes) still launch kernels inside one eager steady-state SMORE step."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

env = bench.build_env("cuda:0", overrides={"cuda_graph": False})
trainer = bench.pkg("trainer").Trainer(env["config"], env["model"])
batches = bench.take_batches(env["train"], 6)
env["model"].train()
for b in batches[:5]:
    trainer._train_batch(b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    trainer._train_batch(batches[5])
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key.startswith("aten::") and e.self_device_time_total > 0:
        rows.append((e.count, e.key, str(e.input_shapes)[:100], e.self_device_time_total))
for c, name, shp, t in sorted(rows, key=lambda r: (-r[0], r[1])):
    print(f"{c:4d}  {t:8.1f} us  {name:26s} {shp}")
