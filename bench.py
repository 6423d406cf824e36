#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200 (see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): SMORE on a synthetic Amazon-Baby-shaped dataset
(19,445 users x 7,050 items x 160,792 interactions, 4096-d image / 384-d text features, d = 64,
train batch 2048, the reference trainer's mirror-gradient schedule). One "step" = one iteration of
Trainer._train_epoch's batch loop: in steady state 3 forward+backward passes and 2 Adam steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* `value`   : training interactions/s over all ranks, device-timed (CUDA events, max over ranks)
              with the K batches already resident in HBM.
* `e2e`     : the same metric through the public API (TrainDataLoader -> Trainer._train_batch):
              host negative sampling, pinned H2D copy of every [3, B] batch and a D2H read of the
              loss inside the timed region.
* extras    : train_epoch_s, eval_users_per_s (Trainer.evaluate on the validation users),
              per-kernel rooflines (SpMM GB/s vs measured HBM peak), cpu_baseline.
* N > 1     : replicas (one hyper-parameter grid point per GPU, the reference's own outer loop);
              no data-path collective, weak scaling.
* --impl reference : the reference's CPU implementation of the same step (oracle port on torch
              CPU with all host threads), rank 0 only.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "recommendar-systems_b200"


def pkg(sub):
    return importlib.import_module(PKG + "." + sub)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_launch"].get(kernel)
    except (OSError, ValueError, KeyError):
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_env(device, seed=999, model_name="SMORE", shape="baby", overrides=None):
    import torch
    synth, cfgm, data_m, models = pkg("synth"), pkg("config"), pkg("data"), pkg("models")
    data = synth.make_dataset(shape)
    cd = {"device": torch.device(device), "data_path": None, "v_feat": data.image_feat,
          "t_feat": data.text_feat, "seed": seed}
    cd.update(overrides or {})
    config = cfgm.Config(model_name, shape, cd)
    ds = data_m.RecDataset(config, data.users, data.items, data.labels)
    tr, va, te = ds.split()
    train = data_m.TrainDataLoader(config, tr, batch_size=config["train_batch_size"], shuffle=True)
    valid = data_m.EvalDataLoader(config, va, additional_dataset=tr, batch_size=config["eval_batch_size"])
    cfgm.init_seed(seed)
    train.pretrain_setup()
    model = models.get_model(model_name)(config, train).to(config["device"])
    return dict(config=config, data=data, tr=tr, train=train, valid=valid, model=model)


def take_batches(loader, n):
    out = []
    while len(out) < n:
        for b in loader:
            out.append(b)
            if len(out) == n:
                loader.pr = 0
                break
    return out


def time_kernel(fn, flush, iters=20, warm=3):
    """Mean device time (ms) of fn() with the L2 flushed before every timed launch."""
    import torch
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters


def kernel_rooflines(env, peak):
    """Per-kernel achieved algorithmic GB/s (SURVEY 8d byte formulas), L2 flushed per launch."""
    import torch
    ops = pkg("ops")
    model, dev = env["model"], env["config"]["device"]
    d = model.embedding_dim
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # 2 x L2
    out = []
    g = model.norm_adj
    X = torch.randn(g.n_cols, d, device=dev)
    Y = torch.empty(g.n_rows, d, device=dev)
    ms = time_kernel(lambda: ops.spmm_raw(g, X, Y=Y), flush)
    b = g.algorithmic_bytes(d)
    out.append({"kernel": "spmm_csr_kernel (UI graph, one layer)", "bytes": b, "ms": ms,
                "achieved": b / ms / 1e6, "frac": b / ms / 1e6 / peak})
    acc = torch.empty_like(X)
    ms = time_kernel(lambda: ops.spmm_raw(g, X, Y=Y, acc_in=X, acc_out=acc), flush)
    b2 = b + 8 * d * g.n_rows
    out.append({"kernel": "spmm_csr_kernel (UI graph, fused layer-sum)", "bytes": b2, "ms": ms,
                "achieved": b2 / ms / 1e6, "frac": b2 / ms / 1e6 / peak})
    gi = model.fusion_adj
    Xi = torch.randn(gi.n_cols, d, device=dev)
    Yi = torch.empty(gi.n_rows, d, device=dev)
    ms = time_kernel(lambda: ops.spmm_raw(gi, Xi, Y=Yi), flush)
    b = gi.algorithmic_bytes(d)
    out.append({"kernel": "spmm_csr_kernel (fusion item-item graph)", "bytes": b, "ms": ms,
                "achieved": b / ms / 1e6, "frac": b / ms / 1e6 / peak})
    # spectral: 2 rows in, 3 rows out
    n = model.n_items
    img, txt = torch.randn(n, d, device=dev), torch.randn(n, d, device=dev)
    w = [model.image_complex_weight[0].detach(), model.text_complex_weight[0].detach(),
         model.fusion_complex_weight[0].detach()]
    ms = time_kernel(lambda: ops.spectrum_convolution(img, txt, *w, True), flush)
    b = 4 * n * d * 5
    out.append({"kernel": "spectral_fwd_kernel", "bytes": b, "ms": ms, "achieved": b / ms / 1e6,
                "frac": b / ms / 1e6 / peak})
    # the table-sized projection GEMMs (image table 7050 x 4096 fp32 = 115 MB, streamed once)
    table = model.image_embedding.weight.detach()
    Wt, bt = model.image_trs.weight.detach(), model.image_trs.bias.detach()
    n, F = table.shape
    dy = torch.randn(n, d, device=dev)
    b = 4 * (n * F + d * F + n * d)
    for name, fn in (("gemm_tc05_kernel (image projection forward)", lambda: ops.gemm(table, True, Wt, True, n, d, F, bt)),
                     ("gemm_tc05_kernel (image projection dW)", lambda: ops.gemm(dy, False, table, False, d, F, n)),
                     ("gemm_tc05_kernel (image projection dX)", lambda: ops.gemm(dy, True, Wt, False, n, F, d))):
        ms = time_kernel(fn, flush)
        out.append({"kernel": name, "bytes": b, "ms": ms, "achieved": b / ms / 1e6, "frac": b / ms / 1e6 / peak})
    # the optimizer pass: p, g, m, v read + p, m, v written = 28 bytes per parameter
    opt = pkg("optim").FusedAdam([torch.nn.Parameter(torch.randn(n, F, device=dev))], lr=1e-3)
    opt.param_groups[0]["params"][0].grad = torch.randn(n, F, device=dev)
    opt.step()
    ms = time_kernel(lambda: opt.step(), flush)
    b = 28 * n * F
    out.append({"kernel": "adam_kernel (image table)", "bytes": b, "ms": ms, "achieved": b / ms / 1e6,
                "frac": b / ms / 1e6 / peak})
    del opt, dy
    # a user-item graph whose embedding table (184 MB) does not fit the 126 MB L2
    del X, Y, acc
    su, si = pkg("synth").make_scaled_edges(dev, 600_000, 120_000, 30_000_000)
    gl = pkg("graph").build_ui_graph(su, si, 600_000, 120_000, "f64eps")
    del su, si
    X = torch.randn(gl.n_cols, d, device=dev)
    Y = torch.empty(gl.n_rows, d, device=dev)
    ms = time_kernel(lambda: ops.spmm_raw(gl, X, Y=Y), flush, iters=10)
    b = gl.algorithmic_bytes(d)
    gather = gl.nnz * (8 + 4 * d) + 4 * (gl.n_rows + 1) + 4 * d * gl.n_rows
    out.append({"kernel": f"spmm_csr_kernel (scaled UI graph 600k x 120k, nnz {gl.nnz}, > L2)", "bytes": b,
                "ms": ms, "achieved": b / ms / 1e6, "frac": b / ms / 1e6 / peak,
                "no_reuse_gather_bytes": gather, "no_reuse_gather_gbps": gather / ms / 1e6})
    for o in out:
        o["unit"] = "GB/s"
    return out


def in_step_kernel_times(trainer, batches, patterns):
    """Average device duration (ms) and launches per step of the kernels whose name contains one
    of `patterns`, measured with the profiler's CUDA activity records (CUPTI) over extra replays of
    the SAME captured step the timed region ran: the live in-step figure, L2 state and co-running
    kernels included."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    try:
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for b in batches:
                trainer._train_batch_graphed(b)
            torch.cuda.synchronize()
        out = {}
        for pat in patterns:
            evs = [e for e in prof.key_averages() if pat in e.key and e.device_time_total > 0]
            n = sum(e.count for e in evs)
            if n:
                out[pat] = {"ms": sum(e.device_time_total for e in evs) / n / 1e3, "per_step": n / len(batches)}
        return out
    except Exception as exc:                    # profiler unavailable: the isolated figures stand
        sys.stderr.write(f"in-step kernel timing skipped: {exc}\n")
        return {}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner off it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lib = pkg("lib")
    lib.load()
    trainer_m = pkg("trainer")
    env = build_env(dev, seed=999 + rank)                # one grid point (seed) per replica
    model, config = env["model"], env["config"]
    trainer = trainer_m.Trainer(config, model)
    B = config["train_batch_size"]
    K, W = args.steps, args.warmup
    batches = take_batches(env["train"], W + K)
    model.train()
    # clocks / throttle reasons are sampled every 50 ms from before the warm-up until the end of the
    # end-to-end loop (the device-timed K steps alone last a few tens of ms: too short for nvidia-smi)
    sampler = ClockSampler(local)
    sampler.start()
    # The steady-state step is replayed from a CUDA graph that is captured on the 5th step (two eager
    # runs per control-flow variant first): with W < 6 a few extra untimed steps go in front, so that
    # the K timed steps are the steady state whatever W the caller chose.
    for b in take_batches(env["train"], max(0, 6 - W)) + batches[:W]:
        trainer._train_batch_graphed(b)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    l0 = lib.launch_count() + trainer.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed_steps")          # ncu --nvtx --nvtx-include "timed_steps/"
    e0.record()
    for b in batches[W:]:
        trainer._train_batch_graphed(b)
    e1.record()
    torch.cuda.nvtx.range_pop()
    barrier()
    launches = lib.launch_count() + trainer.replayed_launches - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * K * B / (ms / 1e3)

    # ---- end to end through the public API: loader (host sampling, pinned H2D) + loss D2H
    # Trainer._train_epoch with the reference's per-batch `loss.item()` (sync_free off): every step
    # draws its batch in the loader, copies it from pinned memory and reads its loss back
    trainer.sync_free = False
    reps = []
    for _ in range(3):                      # K steps three times, median: one host hiccup in a
        barrier()                           # ~60 ms window would otherwise decide the number
        t0 = time.perf_counter()
        done = 0
        while done < K:
            _, lb = trainer._train_epoch(env["train"], 0, max_batches=K - done)
            done += len(lb)
        barrier()
        reps.append(time.perf_counter() - t0)
    e2e_s = sorted(reps)[1]
    trainer.sync_free = True
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * K * B / e2e_s

    # ---- full-rank evaluation (valid users): e2e (incl. D2H + host metrics) and device only
    n_eval = int(env["valid"].eval_u.shape[0])
    for _ in range(3):                       # eager pass, capture pass, first replay
        trainer.evaluate(env["valid"])
    barrier()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        metrics = trainer.evaluate(env["valid"])          # Trainer.evaluate: result dict on the host
    torch.cuda.synchronize()
    eval_s = (time.perf_counter() - t0) / reps
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed_eval")
    a.record()
    trainer.evaluate(env["valid"])
    b_.record()
    torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    eval_dev_ms = a.elapsed_time(b_)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    live = in_step_kernel_times(trainer, batches[W:W + min(K, 5)], ["spmm_csr_kernel<", "adam_kernel", "gemm_tc05_kernel"])
    kr = kernel_rooflines(env, peak)
    dom = dict(kr[1])   # the launch propagate_mean issues: SpMM + fused layer sum over the UI graph
    dom["isolated_ms"], dom["isolated_frac"] = dom["ms"], dom["frac"]
    if "spmm_csr_kernel<" in live:
        # every UI-graph launch of a training step carries the layer-sum epilogue: same bytes
        dom["ms"] = live["spmm_csr_kernel<"]["ms"]
        dom["achieved"] = dom["bytes"] / dom["ms"] / 1e6
        dom["frac"] = dom["achieved"] / peak
        dom["launches_per_step"] = live["spmm_csr_kernel<"]["per_step"]
    n_train = len(env["tr"])
    steps_per_epoch = -(-n_train // B)
    line = {
        "metric": "train interactions/s (SMORE, Baby-shaped; epoch s and eval users/s in extras)",
        "value": value, "unit": "interactions/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SMORE, synthetic Amazon-Baby-shaped (19445 users x 7050 items x 160792 "
                               "interactions, 4096-d image / 384-d text, d=64, batch 2048, mirror-gradient "
                               "schedule: steady-state step = 3 fwd+bwd + 2 Adam)",
                   "l2": "per-step working set (~1.7 GB of tables/optimizer state) exceeds the 126 MB L2; "
                         "kernel rooflines flush L2 before every timed launch",
                   "multi_gpu": "replicas (one seed per GPU), no collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "interactions/s", "h2d_bytes_per_step": 3 * B * 8,
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": dom["frac"], "traffic": ncu_traffic(dom["kernel"]),
                     "traffic_source": "ncu --set full capture committed under profiles/ (r01_ncu_traffic.json)",
                     "algorithmic_bytes_per_launch": dom["bytes"], "ms_per_launch": dom["ms"],
                     "timing": "average duration of the kernel's launches inside replays of the timed step "
                               "(CUPTI activity records); isolated_* = one launch after an L2 flush, CUDA events",
                     "launches_per_step": dom.get("launches_per_step"),
                     "isolated_ms_per_launch": dom["isolated_ms"], "isolated_frac": dom["isolated_frac"],
                     "in_step": live},
        "kernels": kr,
        "train_epoch_s": steps_per_epoch * ms / K / 1e3,
        "train_epoch_s_e2e": steps_per_epoch * e2e_s / K,
        "eval_users_per_s": n_eval / eval_s, "eval_users_per_s_device": n_eval / (eval_dev_ms / 1e3),
        "eval_users": n_eval, "eval_recall@20": metrics.get("recall@20"),
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(env, steps=20, warmup=3)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def oracle_trainer_for(env):
    """The reference's CPU path for the same step: oracle port on torch CPU."""
    import torch
    from oracle import build as obuild
    from oracle.train import OracleTrainer
    data, tr, config = env["data"], env["tr"], env["config"]
    cfg = {k: config[k] for k in ("n_ui_layers", "n_layers", "image_knn_k", "text_knn_k", "reg_weight",
                                  "cl_loss", "train_batch_size")}
    G, _ = obuild.build_graphs("SMORE", tr.users, tr.items, data.n_users, data.n_items, data.image_feat,
                               data.text_feat, cfg)
    P = {k: v.detach().cpu() for k, v in env["model"].state_dict().items()}
    drop = torch.nn.Dropout(p=config["dropout_rate"])
    return OracleTrainer("SMORE", P, G, cfg, lr=config["learning_rate"],
                         lr_scheduler=tuple(config["learning_rate_scheduler"]), dropout=drop)


def cpu_baseline(env, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count())
    ot = oracle_trainer_for(env)
    B = env["config"]["train_batch_size"]
    batches = [b.cpu() for b in take_batches(env["train"], warmup + steps)]
    for b in batches[:warmup]:
        ot.step(b)
    t0 = time.perf_counter()
    for b in batches[warmup:]:
        ot.step(b)
    dt = time.perf_counter() - t0
    return {"value": steps * B / dt, "unit": "interactions/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steady-state (mirror-gradient) SMORE training steps of the same workload "
                      f"after {warmup} warm-up steps, oracle port on torch CPU; {dt / steps:.2f} s/step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count())
    dev = "cuda:0" if torch.cuda.is_available() else None
    if dev is None:
        print(json.dumps({"impl": "reference", "unavailable": "model/graph setup needs the CUDA library"}))
        return
    env = build_env(dev)
    K, W = args.steps, args.warmup
    B = env["config"]["train_batch_size"]
    ot = oracle_trainer_for(env)
    batches = [b.cpu() for b in take_batches(env["train"], W + K)]
    for b in batches[:W]:
        ot.step(b)
    t0 = time.perf_counter()
    for b in batches[W:]:
        ot.step(b)
    dt = time.perf_counter() - t0
    value = K * B / dt
    n_train = len(env["tr"])
    line = {"impl": "reference",
            "metric": "train interactions/s (SMORE, Baby-shaped; epoch s and eval users/s in extras)",
            "value": value, "unit": "interactions/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SMORE, synthetic Amazon-Baby-shaped, reference CPU path (oracle port)"},
            "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": torch.get_num_threads(),
                             "kind": "port", "sample": f"{K} training steps after {W} warm-up steps"},
            "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "train_epoch_s": -(-n_train // B) * dt / K}
    print(json.dumps(line))


def run_scaled(args):
    """Config 5: scaled power-law graph, LightGCN-style propagation (row-sharded, NCCL all-gather per
    layer) + item-sharded full-rank top-50 (local fused top-K + all-gather + merge). Strong scaling:
    every rank works on the same users. `--scale 1.0` = 10M users x 2M items x 500M interactions."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(dev))
    ops, G, par, synth, lib = pkg("ops"), pkg("graph"), pkg("parallel"), pkg("synth"), pkg("lib")
    U, I, E = int(10_000_000 * args.scale), int(2_000_000 * args.scale), int(500_000_000 * args.scale)
    d, L, k, Bu = 64, 4, 50, args.eval_users
    su, si = synth.make_scaled_edges(dev, U, I, E)
    full = G.build_ui_graph(su, si, U, I, "f64eps")
    del su, si
    gen = torch.Generator(device=dev).manual_seed(999)
    bound = (6.0 / (U + I + d)) ** 0.5
    X0 = (torch.rand(U + I, d, generator=gen, device=dev) * 2 - 1) * bound
    users = torch.randint(0, U, (Bu,), generator=gen, device=dev)
    nnz_full = full.nnz
    if world > 1 and args.partition == "rows":
        sg = par.ShardedUIGraph(full, rank, world)
        lo, hi = par.item_range(I, rank, world)
        del full
    elif world > 1:
        sb = par.ShardedBipartite(full, rank, world)
        Xu, Xi = X0[sb.lo: sb.hi].contiguous(), X0[U:].contiguous()
        del X0
    torch.cuda.empty_cache()

    def step():
        if world == 1:
            out = ops.propagate_mean(full, X0, L)
            return ops.score_mask_topk(out[:U], users, out[U:], k)
        if args.partition == "rows":
            out = par.sharded_propagate_mean(sg, X0, L)
            return par.sharded_score_topk(out[:U], users, out[U + lo: U + hi].contiguous(), lo, k)
        ou, oi = par.bipartite_propagate_mean(sb, Xu, Xi, L)
        return par.bipartite_score_topk(sb, ou, oi, users, k)

    with torch.no_grad():
        sampler = ClockSampler(local)     # from before the warm-up: the timed steps alone are < 100 ms at 8 GPUs
        sampler.start()
        for _ in range(args.warmup):
            ids = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ids = step()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches = lib.launch_count() - l0
        clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        chk = torch.tensor([float(ids.sum().item())], device=dev, dtype=torch.float64)
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        assert lo_.item() == hi_.item(), "ranks disagree on the merged top-K"
    if rank == 0:
        peak, peak_src = measured_peaks()
        nnz = nnz_full if world == 1 else None
        n = U + I
        line = {"metric": "propagate (4 layers) + full-rank top-50, eval users/s (scaled power-law graph)",
                "value": Bu * args.steps / (ms / 1e3), "unit": "users/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"LightGCN-style propagation + item-sharded top-50, {U} users x {I} items x "
                                       f"~{E} interactions, d=64, {Bu} eval users per step",
                           "partition": ("rows by nnz + all-gather of the whole table per layer" if args.partition == "rows"
                                         else "users by nnz, items replicated: one all-reduce of the [I, d] item table per "
                                              "layer, overlapped with the user-side SpMM") + "; evaluation: items by range + "
                                                                                             "top-K merge",
                           "l2": "embedding table larger than L2" if n * d * 4 > 126e6 else "fits L2"},
                "clocks": clocks, "gpu_launches": int(launches), "checksum_ids": int(ids.sum().item())}
        if nnz is not None:
            b = L * (8 * nnz + 4 * (n + 1) + 8 * d * n)
            line["roofline"] = {"bound": "hbm", "kernel": "spmm_csr_kernel x4 (propagation share of the step)",
                                "algorithmic_bytes_per_step": b, "peak": peak, "peak_source": peak_src, "unit": "GB/s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="smore_baby", choices=["smore_baby", "scaled"])
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--eval-users", type=int, default=16384)
    ap.add_argument("--partition", default="bipartite", choices=["bipartite", "rows"])
    args = ap.parse_args()
    if args.workload == "scaled" and args.impl == "ours":
        run_scaled(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
