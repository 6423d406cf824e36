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

# ---- how much of the step is the GPU idle? (union of the kernel intervals of all streams against the span) ----
ev = [(e.time_range.start, e.time_range.end) for e in prof.events()
      if e.device_type.name == "CUDA" and e.time_range.end > e.time_range.start]
ev.sort()
if ev:
    span = ev[-1][1] - ev[0][0]
    busy, cur_s, cur_e, gaps = 0.0, ev[0][0], ev[0][1], []
    for s_, e_ in ev[1:]:
        if s_ > cur_e:
            busy += cur_e - cur_s
            gaps.append(s_ - cur_e)
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    busy += cur_e - cur_s
    big = [g for g in gaps if g > 20]                 # step boundaries (host replay latency), not kernel gaps
    small = [g for g in gaps if g <= 20]
    print(f"# timeline: span {span / steps:.1f} us/step, GPU busy (any stream) {busy / steps:.1f} us/step, "
          f"idle {sum(gaps) / steps:.1f} us/step = {len(small) / steps:.0f} gaps <= 20 us totalling {sum(small) / steps:.1f} us "
          f"(median {sorted(small)[len(small) // 2] if small else 0:.2f} us) + {len(big) / steps:.1f} gaps > 20 us totalling {sum(big) / steps:.1f} us")
