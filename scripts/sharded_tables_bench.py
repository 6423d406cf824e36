"""Item-range sharded modality tables under torchrun: SMORE training step with
the [I, 4096] / [I, 384] tables, their projection GEMMs, gradients and Adam state split over the
ranks (parallel.sharded_projection), everything else replicated.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/sharded_tables_bench.py [shape] [d] [--check]

Prints one JSON line on rank 0: device ms/step (max over ranks). --check also trains an
unsharded replica on every rank and compares the loss trajectory and the local table rows."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
shape = argv[0] if argv else "clothing"
d = int(argv[1]) if len(argv) > 1 else 128
check = "--check" in sys.argv
K, W = 20, 8
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))


def run(shard):
    over = {"embedding_size": d}
    if shard:
        over["table_shard"] = (rank, world)
    env = bench.build_env(dev, model_name="SMORE", shape=shape, overrides=over)
    trainer = bench.pkg("trainer").Trainer(env["config"], env["model"])
    batches = bench.take_batches(env["train"], W + K)
    env["model"].train()
    env["model"].pre_epoch_processing()
    losses = []
    for b in batches[:W]:
        losses.append(trainer._train_batch_graphed(b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for b in batches[W:]:
        losses.append(trainer._train_batch_graphed(b))
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(e) / K], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), torch.stack(losses).float().cpu(), env


ms, losses, env = run(shard=world > 1)
out = {"workload": f"SMORE {shape} d={d}, item-range sharded feature tables", "n_gpus": world,
       "ms_per_step": round(ms, 3), "interactions_per_s": round(env["config"]["train_batch_size"] / ms * 1e3)}
if check:
    m = env["model"]
    rows = m._table_rows
    img = m.image_embedding.weight.detach().clone()
    del env
    torch.cuda.empty_cache()
    ms1, losses1, env1 = run(shard=False)
    ref = env1["model"].image_embedding.weight.detach()
    ref = ref if rows is None else ref[rows.lo: rows.hi]
    out["ms_per_step_unsharded"] = round(ms1, 3)
    out["max_rel_loss_diff"] = float(((losses - losses1).abs() / losses1.abs()).max())
    out["table_rel_diff"] = float((img - ref).abs().max() / ref.abs().max())
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
