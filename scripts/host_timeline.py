"""Where the wall clock of an end-to-end epoch goes (Trainer._train_epoch, per-batch loss read): the same loop
instrumented piece by piece -- enqueue (static copy + graph replay), loader.next (host negative sampling + pinned
H2D), loss read-back wait -- plus the epoch boundary (iterator set-up, shuffle)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
env = bench.build_env("cuda:0")
tr = bench.pkg("trainer").Trainer(env["config"], env["model"])
tr.sync_free = False
for _ in range(3):
    tr._train_epoch(env["train"], 0)
torch.cuda.synchronize()
pc = time.perf_counter
T = {"iter": 0.0, "enqueue": 0.0, "next": 0.0, "item": 0.0, "shuffle": 0.0}
n_steps = 0
t_all = pc()
for ep in range(3):
    t = pc(); it = iter(env["train"]); inter = next(it, None); T["iter"] += pc() - t
    unread = []
    while inter is not None:
        t = pc(); loss = tr._train_batch_graphed(inter, n_steps); T["enqueue"] += pc() - t
        n_steps += 1
        t = pc(); inter = next(it, None); T["next"] += pc() - t
        if inter is None and hasattr(env["train"], "prefetch_shuffle"):
            t = pc(); env["train"].prefetch_shuffle(); T["shuffle"] += pc() - t
        unread.append(loss)
        t = pc()
        while len(unread) > (1 if inter is not None else 0):
            unread.pop(0).item()
        T["item"] += pc() - t
torch.cuda.synchronize()
wall = pc() - t_all
print(f"3 epochs, {n_steps} steps: {wall / 3 * 1e3:.2f} ms per epoch, {wall / n_steps * 1e3:.3f} ms per step")
for k, v in T.items():
    print(f"  {k:8s} {v / 3 * 1e3:8.2f} ms per epoch   {v / n_steps * 1e6:8.1f} us per step")
