"""GPU tests added in the second half of round 2 (all through the C ABI):
* nn.Dropout of the SMORE preference module generated inside the kernels (smore.py:331-333): the
  generator against oracle/dropout.py bit for bit, the fused modules with in-kernel dropout against
  the same modules fed the materialised masks (bit-identical), fresh masks on CUDA-graph replays.
"""
import numpy as np
import pytest
import torch

from conftest import pkg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _counter(v):
    return torch.tensor([v], dtype=torch.float64, device=DEV)


@pytest.mark.parametrize("planes,n,d,p,seed,count", [(3, 26495, 64, 0.5, 999, 0), (3, 1, 32, 0.1, (999 << 32) ^ 12345, 17),
                                                    (1, 777, 128, 0.25, 2 ** 64 - 1, 2 ** 31 + 5), (3, 100, 64, 0.0, 1, 1)])
def test_inkernel_dropout_generator_matches_oracle(planes, n, d, p, seed, count):
    from oracle import dropout as odrop
    ops = pkg("ops")
    got = ops.dropout_mask(planes, n, d, (p, seed, _counter(count)))
    want = odrop.dropout_multipliers(planes, n, d, p, seed, count)
    assert np.array_equal(got.cpu().numpy(), want)
    # counter = NULL reads as 0
    got0 = ops.dropout_mask(planes, n, d, (p, seed, None), device=torch.device(DEV))
    assert np.array_equal(got0.cpu().numpy(), odrop.dropout_multipliers(planes, n, d, p, seed, 0))


def _layers(d, seed):
    torch.manual_seed(seed)
    mk = lambda bias: torch.nn.Linear(d, d, bias=bias).to(DEV)
    return [mk(True), mk(False), mk(True), mk(False), mk(True), mk(True), mk(True)]


@pytest.mark.parametrize("n,d,p", [(26495, 64, 0.5), (333, 32, 0.2), (6001, 128, 0.1), (65, 64, 0.9)])
def test_smore_side_inkernel_dropout_equals_explicit_masks(n, d, p, monkeypatch):
    """mmrec_smore_side_*_drop_f32 == mmrec_smore_side_*_f32 fed the masks the generator writes out
    (forward outputs and every gradient bit-identical): forward and backward regenerate one mask.
    (Both on the mma.sync forward: the tcgen05 forward is compared with it in its own test.)"""
    monkeypatch.setenv("MMREC_SIDE_TC", "0")
    ops = pkg("ops")
    layers = _layers(d, 5)
    gen = torch.Generator().manual_seed(n)
    ins = [torch.randn(n, d, generator=gen).to(DEV) for _ in range(4)]
    ga, gs = torch.randn(n, d, generator=gen).to(DEV), torch.randn(n, d, generator=gen).to(DEV)
    spec = (p, (999 << 32) ^ 77, _counter(41))
    res = []
    for mode in ("drop", "masks"):
        x = [t.clone().requires_grad_(True) for t in ins]
        for l in layers:
            l.zero_grad()
        if mode == "drop":
            a, s = ops.smore_side(*x, layers, None, spec)
        else:
            a, s = ops.smore_side(*x, layers, ops.dropout_mask(3, n, d, spec))
        ((a * ga).sum() + (s * gs).sum()).backward()
        res.append([a.detach(), s.detach()] + [t.grad for t in x] + [l.weight.grad.clone() for l in layers] +
                   [l.bias.grad.clone() for l in layers if l.bias is not None])
    for u, v in zip(*res):
        assert torch.equal(u, v)
    # and the masks matter: without dropout the output differs
    a0, _ = ops.smore_side(*ins, layers)
    assert not torch.equal(a0, res[0][0])


@pytest.mark.parametrize("n,d,p", [(62420, 128, 0.5), (777, 64, 0.2), (33, 32, 0.1)])
def test_smore_combine_inkernel_dropout_equals_explicit_masks(n, d, p):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(n + d)
    names = ("zv", "zt", "V", "T", "F", "C", "gi", "gt", "gf")
    t = {k: torch.randn(n, d, generator=gen).to(DEV) for k in names}
    for k in ("gi", "gt", "gf"):
        t[k] = torch.sigmoid(t[k])
    ga, gs = torch.randn(n, d, generator=gen).to(DEV), torch.randn(n, d, generator=gen).to(DEV)
    spec = (p, 31337, _counter(3))
    res = []
    for mode in ("drop", "masks"):
        y = {k: v.clone().requires_grad_(True) for k, v in t.items()}
        args = [y[k] for k in names]
        if mode == "drop":
            a, s = ops.smore_combine(*args, None, spec)
        else:
            a, s = ops.smore_combine(*args, ops.dropout_mask(3, n, d, spec))
        ((a * ga).sum() + (s * gs).sum()).backward()
        res.append([a.detach(), s.detach()] + [y[k].grad for k in names])
    for u, v in zip(*res):
        assert torch.equal(u, v)


@pytest.mark.parametrize("tc", ["0", "1"])
def test_inkernel_dropout_draws_fresh_masks_on_graph_replay(tc, monkeypatch):
    """A captured launch reads the counter on the device: replays after the counter moved use the
    mask of the new count (what FusedAdam's update count provides inside a captured training step).
    Both forwards: mma.sync (bit-identical to the explicit-mask module) and tcgen05 (to fp32 rounding)."""
    monkeypatch.setenv("MMREC_SIDE_TC", tc)
    ops = pkg("ops")
    n, d, p = 4000, 64, 0.5
    layers = _layers(d, 9)
    ins = [torch.randn(n, d, device=DEV) for _ in range(4)]
    cnt = _counter(5)
    spec = (p, 4242, cnt)
    with torch.no_grad():
        ops.smore_side(*ins, layers, None, spec)                  # warm-up outside capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            a, s = ops.smore_side(*ins, layers, None, spec)
        outs = []
        for c in (5, 6, 6, 1000):
            cnt.fill_(c)
            g.replay()
            want, _ = ops.smore_side(*ins, layers, ops.dropout_mask(3, n, d, (p, 4242, _counter(c))))
            if tc == "0":
                assert torch.equal(a, want)
            else:
                assert float((a - want).abs().max() / want.abs().max()) < 2e-6
            outs.append(a.clone())
    assert not torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2]) and not torch.equal(outs[2], outs[3])


def test_smore_trainer_uses_inkernel_dropout_under_graph_replay():
    """SMORE with dropout_rate > 0 through the Trainer: no mask tensors are drawn (torch's CUDA
    generator is not consumed by the step), the dropout counter is the optimizer's device-side
    update count, and graph-replayed steps stay finite and keep training."""
    from parity_util import make_env
    env = make_env("SMORE", DEV, overrides={"cuda_graph": True, "dropout_rate": 0.5})
    m = env["model"]
    tr = pkg("trainer").Trainer(env["config"], m)
    assert m.fused_dropout and m.dropout_counter is not None
    assert m.dropout_counter.data_ptr() == tr.optimizer.param_groups[0]["hyper"].data_ptr() + 8
    state = torch.cuda.get_rng_state(0)
    losses = []
    for epoch in range(3):
        total, _ = tr._train_epoch(env["train"], epoch)
        losses.append(float(total))
    assert torch.equal(state, torch.cuda.get_rng_state(0))        # nn.Dropout no longer touches torch's generator
    assert tr.replayed_launches > 0 and np.isfinite(losses).all()
    assert float(m.dropout_counter.item()) > 0


# ------------------------------------------------------------------ activations on the special-function unit
def test_fast_activations():
    """tanh / sigmoid / exp as the fused kernels evaluate them (ex2.approx / rcp.approx + a polynomial
    for small |x|, csrc/common.cuh) against float64: relative error far below the 1e-5 parity bar over
    the whole range, exact limits, no NaN at the extremes."""
    lib = pkg("lib")
    xs = torch.cat([torch.linspace(-30, 30, 2_000_001), torch.logspace(-30, 1.5, 200_001), -torch.logspace(-30, 1.5, 200_001),
                    torch.tensor([0.0, -0.0, 0.6, -0.6, 0.59999996, 88.0, -88.0, 1e4, -1e4, 1e-38, -1e-38])]).to(DEV)
    x64 = xs.double()
    want = {1: torch.tanh(x64), 2: torch.sigmoid(x64), 3: torch.exp(x64)}
    bound = {1: 4e-7, 2: 1e-6, 3: 2e-6}          # exp: the argument rounding grows with |x| (softmax only sees x <= 0)
    for act in (1, 2, 3):
        y = torch.empty_like(xs)
        lib.call("mmrec_activation_f32", lib.ptr(xs), xs.numel(), act, lib.ptr(y), lib.stream())
        assert bool(torch.isfinite(y[xs.abs() < 80]).all())
        ok = want[act].abs() > 1e-30
        if act == 3:
            ok &= xs.abs() < 30
        if act == 2:
            # far in the negative tail sigmoid(x) ~ e^x inherits exp's argument rounding (|x| * 6e-8): relative
            # to values of 1e-13 that is irrelevant; what matters is the absolute error on the (0, 1) scale
            assert float((y.double() - want[act]).abs().max()) < 1.5e-7
            ok &= xs.abs() <= 8
        rel = ((y.double() - want[act]).abs() / want[act].abs())[ok]
        assert float(rel.max()) < bound[act], (act, float(rel.max()))
    t = torch.empty_like(xs)
    lib.call("mmrec_activation_f32", lib.ptr(xs), xs.numel(), 1, lib.ptr(t), lib.stream())
    assert float(t[xs == 1e4]) == 1.0 and float(t[xs == -1e4]) == -1.0 and float(t[xs == 0][0]) == 0.0
    assert bool((t.abs() <= 1).all())


def test_smore_side_inference_forward_keeps_nothing():
    """Under no_grad (full_sort_predict) the fused preference module writes no saved tensors
    (saved = NULL in mmrec_smore_side_fwd_f32) and returns the same bits as the training forward."""
    ops = pkg("ops")
    n, d = 7050 + 19445, 64
    layers = _layers(d, 3)
    ins = [torch.randn(n, d, device=DEV) for _ in range(4)]
    a1, s1 = ops.smore_side(*ins, layers)
    before = torch.cuda.memory_allocated()
    with torch.no_grad():
        a0, s0 = ops.smore_side(*ins, layers)
    assert torch.equal(a0, a1) and torch.equal(s0, s1)
    assert not a0.requires_grad and a1.requires_grad
    assert torch.cuda.memory_allocated() - before <= 2 * n * d * 4 + (1 << 20)      # the two outputs, no [7, n, d]


# ------------------------------------------------------------------ a10: preference module forward on tcgen05
@pytest.mark.parametrize("n,p,keep", [(26495, 0.5, True), (26495, 0.0, False), (1, 0.0, True), (129, 0.3, True),
                                      (128 * 148 * 2 + 77, 0.1, True)])
def test_smore_side_forward_tcgen05_matches_mma_sync_and_float64(n, p, keep, monkeypatch):
    """mmrec_smore_side_fwd_tc_f32 (d = 64: A operands in TMEM, pre-split weight images streamed by bulk
    copies, sixteen epilogue warps) against the mma.sync forward and against float64: outputs and the
    seven saved tensors to fp32 rounding, the same dropout pattern, the mma.sync backward on its saved
    tensors gives the same gradients; ragged last tile, one row, more tiles than SMs, inference
    (saved = NULL)."""
    ops = pkg("ops")
    d = 64
    layers = _layers(d, 11)
    for l in layers:
        l.weight.data *= 3.0
    gen = torch.Generator().manual_seed(n)
    ins = [torch.randn(n, d, generator=gen).to(DEV) for _ in range(4)]
    ga, gs = torch.randn(n, d, generator=gen).to(DEV), torch.randn(n, d, generator=gen).to(DEV)
    spec = (p, 99, _counter(7)) if p > 0 else None
    res = {}
    for tc in ("0", "1"):
        monkeypatch.setenv("MMREC_SIDE_TC", tc)
        x = [t.clone().requires_grad_(keep) for t in ins]
        for l in layers:
            l.zero_grad()
        if keep:
            a, s = ops.smore_side(*x, layers, None, spec)
            ((a * ga).sum() + (s * gs).sum()).backward()
            res[tc] = [a.detach(), s.detach()] + [t.grad for t in x] + [l.weight.grad.clone() for l in layers]
        else:
            with torch.no_grad():
                a, s = ops.smore_side(*x, layers, None, spec)
            res[tc] = [a, s]
    rel = lambda u, v: float((u.double() - v.double()).abs().max() / v.double().abs().max().clamp_min(1e-12))
    for i, (u, v) in enumerate(zip(res["1"], res["0"])):
        assert rel(u, v) < 5e-6, i
    if p > 0:      # the same elements are dropped: side - (pf f)/3 ... is not observable directly; compare zero patterns of a probe
        monkeypatch.setenv("MMREC_SIDE_TC", "1")
        probe = [torch.ones(n, d, device=DEV) for _ in range(4)]
        a1, s1 = ops.smore_side(*probe, layers, None, spec)
        monkeypatch.setenv("MMREC_SIDE_TC", "0")
        a0, s0 = ops.smore_side(*probe, layers, None, spec)
        assert rel(s1, s0) < 5e-6
    # float64 reference (masks materialised by the generator)
    masks = ops.dropout_mask(3, n, d, spec).double().cpu() if p > 0 else None
    lin = torch.nn.functional.linear
    L64 = [torch.nn.Linear(d, d, bias=l.bias is not None).double() for l in layers]
    for a_, b_ in zip(L64, layers):
        a_.load_state_dict({k: v.double().cpu() for k, v in b_.state_dict().items()})
    F, V, T, C = [t.double().cpu() for t in ins]
    qv = lin(torch.tanh(lin(F, L64[0].weight, L64[0].bias)), L64[1].weight)
    qt = lin(torch.tanh(lin(F, L64[2].weight, L64[2].bias)), L64[3].weight)
    gate = lambda l: torch.sigmoid(lin(C, l.weight, l.bias))
    m = masks if masks is not None else torch.ones(3, n, d, dtype=torch.float64)
    side64 = (m[0] * gate(L64[4]) * torch.softmax(qv, -1) * V + m[1] * gate(L64[5]) * torch.softmax(qt, -1) * T +
              m[2] * gate(L64[6]) * F) / 3
    assert rel(res["1"][1].cpu(), side64) < 1e-5 and rel(res["1"][0].cpu(), C + side64) < 1e-5


# ------------------------------------------------------------------ a6: programmatic dependent launch (opt-in)
def test_spmm_chain_under_programmatic_dependent_launch_is_bit_identical(monkeypatch):
    """MMREC_PDL=1: the SpMM launches of a propagation carry the programmatic-serialization attribute,
    read their task / index round before griddepcontrol.wait and everything the previous launch wrote
    after it. Eager and CUDA-graph replay, forward chain and Horner backward: same bits as MMREC_PDL=0."""
    ops, G, synth = pkg("ops"), pkg("graph"), pkg("synth")
    d = synth.make_dataset("tiny", features=False)
    u, i = d.split(0)
    g = G.build_ui_graph(torch.from_numpy(u).to(DEV), torch.from_numpy(i).to(DEV), d.n_users, d.n_items, "f32")
    X = torch.randn(g.n_cols, 64, device=DEV)
    res = {}
    for pdl in ("0", "1"):
        monkeypatch.setenv("MMREC_PDL", pdl)
        y = ops._PropagateMean.apply(X, g, 4)
        z = ops._horner(g, y, 4, 0.2)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ops._PropagateMean.apply(X, g, 4)
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                yg = ops._PropagateMean.apply(X, g, 4)
                zg = ops._horner(g, yg, 4, 0.2)
            for _ in range(3):
                gr.replay()
            s.synchronize()
        assert torch.equal(y, yg) and torch.equal(z, zg)
        res[pdl] = (y, z)
    assert torch.equal(res["0"][0], res["1"][0]) and torch.equal(res["0"][1], res["1"][1])


# ------------------------------------------------------------------ a10: the preference module on the batch rows
@pytest.mark.parametrize("shape,d,p", [("tiny", 64, 0.0), ("tiny", 64, 0.5), ("small", 32, 0.3), ("small", 128, 0.2),
                                       ("baby", 64, 0.5)])
def test_smore_batch_rows_training_equals_all_rows(shape, d, p):
    """SMORE.calculate_loss with the preference module evaluated on the 3 B rows of the batch
    (ops.gather_batch_rows; smore.py:395-407 consumes nothing else of it) against the same step over all
    rows: the same loss and the same gradient for every parameter -- with dropout too (the compact call
    draws the multipliers of the rows it gathered). Repeated users / items inside the batch included."""
    import bench
    over = {"embedding_size": d, "dropout_rate": p, "cuda_graph": False}
    res = {}
    # all rows (the reference's evaluation) | batch rows gathered from the dense views | batch rows with the
    # user rows R x' of the views formed for the batch users only (ops.gather_batch_views)
    modes = {False: dict(batch_rows=False), "rows": dict(batch_rows=True, batch_views=False),
             True: dict(batch_rows=True, batch_views=True)}
    for mode, flags in modes.items():
        env = bench.build_env(DEV, shape=shape, overrides=dict(over, **flags))
        m = env["model"]
        m.train()
        m.dropout_counter = _counter(5)
        assert m.batch_rows is flags["batch_rows"] and m.batch_views is flags.get("batch_views", False)
        batch = bench.take_batches(env["train"], 1)[0]
        batch[0, 1::7] = batch[0, 0]                    # one user many times
        batch[1, 2::5] = batch[1, 1]                    # one positive item many times
        batch[2, ::3] = batch[1, ::3].roll(1)           # negatives that are other rows' positives
        m.zero_grad()
        loss = m.calculate_loss(batch)
        loss.backward()
        res[mode] = (float(loss.detach()), {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None})
    l0, g0 = res[False]
    for mode in ("rows", True):
        l1, g1 = res[mode]
        assert abs(l0 - l1) <= 2e-6 * abs(l0), mode
        assert set(g0) == set(g1)
        for k in g0:
            scale = float(g0[k].abs().max().clamp_min(1e-12))
            assert float((g0[k] - g1[k]).abs().max()) <= 2e-5 * scale, (mode, k)


@pytest.mark.parametrize("shape", ["tiny", "small", "sports"])
def test_mgcn_batch_rows_training_equals_all_rows(shape):
    """MGCN.calculate_loss with the attention fuser / preference gates evaluated on the batch rows and the user
    rows R x' of the two views formed for the batch users only, against the all-rows step (mgcn.py:146-253):
    same loss, same gradients."""
    import bench
    res = {}
    for mode in (False, True):
        env = bench.build_env(DEV, model_name="MGCN", shape=shape, overrides={"cuda_graph": False, "batch_rows": mode})
        m = env["model"]
        m.train()
        assert m.batch_rows is mode
        batch = bench.take_batches(env["train"], 1)[0]
        batch[0, 1::7] = batch[0, 0]
        batch[1, 2::5] = batch[1, 1]
        batch[2, ::3] = batch[1, ::3].roll(1)
        m.zero_grad()
        loss = m.calculate_loss(batch)
        loss.backward()
        res[mode] = (float(loss.detach()), {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None})
    (l0, g0), (l1, g1) = res[False], res[True]
    assert abs(l0 - l1) <= 2e-6 * abs(l0)
    assert set(g0) == set(g1)
    # query_common.0.bias is a sum that cancels to ~1e-10 (softmax over two logits: d logit_t = -d logit_i, so the
    # bias gradient is sum_rows d logit_i * w2 * (h_t^2 - h_i^2)); its two evaluations differ by 3e-14 absolute =
    # float32 rounding of the terms (scripts/debug_mgcn_batch.py, profiles/r02_mgcn_batch_rows.txt). Hence the
    # absolute floor of 1e-9 of the largest gradient of the model, 60x below float32 epsilon on that scale.
    floor = 1e-9 * max(float(g.abs().max()) for g in g0.values())
    for k in g0:
        scale = float(g0[k].abs().max().clamp_min(1e-12))
        assert float((g0[k] - g1[k]).abs().max()) <= 2e-5 * scale + floor, k
