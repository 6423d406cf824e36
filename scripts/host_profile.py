"""cProfile of the host side of the end-to-end training loop (Trainer._train_epoch, per-batch loss read)."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
env = bench.build_env("cuda:0")
tr = bench.pkg("trainer").Trainer(env["config"], env["model"])
tr.sync_free = False
tr._train_epoch(env["train"], 0, max_batches=10)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
tr._train_epoch(env["train"], 0, max_batches=40)
torch.cuda.synchronize()
pr.disable()
print("ms/step", (time.perf_counter() - t0) / 40 * 1e3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
