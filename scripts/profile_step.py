"""Kernel-level time breakdown of steady-state SMORE training steps (torch.profiler, CUDA only)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

model_name = sys.argv[1] if len(sys.argv) > 1 else "SMORE"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
shape = sys.argv[3] if len(sys.argv) > 3 else "baby"
over = {} if model_name in ("SMORE", "MGCN", "FREEDOM") else {"is_multimodal_model": False}
if len(sys.argv) > 4:
    over["embedding_size"] = int(sys.argv[4])
env = bench.build_env("cuda:0", model_name=model_name, shape=shape, overrides=over)
trainer = bench.pkg("trainer").Trainer(env["config"], env["model"])
batches = bench.take_batches(env["train"], 8 + steps)
env["model"].train()
env["model"].pre_epoch_processing()
for b in batches[:8]:
    trainer._train_batch_graphed(b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for b in batches[8:]:
        trainer._train_batch_graphed(b)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([(e.device_time_total, e.count, e.key) for e in ka if e.device_time_total > 0 and e.device_type.name == "CUDA"],
              reverse=True)
tot = sum(r[0] for r in rows)
print(f"# {model_name}: {steps} steps, total device kernel time {tot / steps / 1e3:.3f} ms/step")
for t, c, k in rows[:45]:
    print(f"{t / steps:10.1f} us/step {100 * t / tot:5.1f}%  x{c / steps:6.1f}  {k[:110]}")
