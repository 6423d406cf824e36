"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic: row partition, padded numbering,
per-layer all-gather, item-sharded top-K + merge. The per-rank arithmetic is injected from the
oracle (the CUDA kernels are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import TINY, pkg
from oracle import graph as ograph
from oracle import ops as oops


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_graph(u, i, U, I):
    G = pkg("graph")
    r, c, v = ograph.norm_adj_f32(u, i, U, I)
    n = U + I
    row_ptr = torch.from_numpy(np.concatenate(([0], np.cumsum(np.bincount(r, minlength=n)))).astype(np.int32))
    return G.CSRGraph(row_ptr, torch.from_numpy(c.astype(np.int32)), torch.from_numpy(v), n, n, symmetric=True), (r, c, v)


def _local_spmm_cpu(sg, x_pad, y_slice, acc_in, acc_out, scale):
    g = sg.local
    rp = g.row_ptr.long()
    A = torch.sparse_csr_tensor(rp, g.col_idx.long(), g.vals.double(), size=(g.n_rows, g.n_cols))
    y = torch.sparse.mm(A, x_pad.double()).to(x_pad.dtype)
    if y_slice is not None:
        y_slice.copy_(y)
    if acc_out is not None:
        acc_out.copy_(((acc_in + y) if acc_in is not None else y) * scale)


def _merge_cpu(vals, idx):
    L, n, k = vals.shape
    v = vals.permute(1, 0, 2).reshape(n, L * k).numpy()
    i = idx.permute(1, 0, 2).reshape(n, L * k).numpy().astype(np.int64)
    order = np.lexsort((i, -v), axis=1)[:, :k]
    return torch.from_numpy(np.take_along_axis(i, order, 1))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        par, synth = pkg("parallel"), pkg("synth")
        data = synth.make_dataset("tiny", **TINY)
        u, i = data.split(0)
        U, I = data.n_users, data.n_items
        full, (r, c, v) = _cpu_graph(u, i, U, I)
        sg = par.ShardedUIGraph(full, rank, world)
        assert sg.layout.bounds[0] == 0 and sg.layout.bounds[-1] == U + I
        # padded numbering round trip
        X = torch.randn(U + I, 8, generator=torch.Generator().manual_seed(1))
        assert torch.equal(sg.layout.unpad(sg.layout.pad(X)), X)
        ids = torch.arange(U + I)
        pid = sg.layout.to_padded(ids)
        assert torch.equal(sg.layout.pad(X)[pid], X)
        # sharded propagation == single-process oracle, forward and backward
        A = ograph.to_torch_csr(r, c, v, (U + I, U + I), torch.float64)
        Xd = X.double().requires_grad_(True)
        want = oops.propagate_mean(A, Xd, 3)
        Xs = X.double().requires_grad_(True)
        got = par.sharded_propagate_mean(sg, Xs, 3, local_spmm=_local_spmm_cpu)
        assert torch.allclose(got, want, atol=1e-12)
        W = torch.randn(U + I, 8, generator=torch.Generator().manual_seed(2)).double()
        (want * W).sum().backward()
        (got * W).sum().backward()
        assert torch.allclose(Xs.grad, Xd.grad, atol=1e-12)
        # item-sharded top-K + merge == global stable sort
        ue = torch.randn(50, 16, generator=torch.Generator().manual_seed(3))
        ie = torch.randn(I, 16, generator=torch.Generator().manual_seed(4))
        ie[7] = ie[2]                                           # exact tie across / inside shards
        users = torch.arange(50)
        lo, hi = par.item_range(I, rank, world)

        def local_topk(ue_, users_, ie_loc, item_lo, k, mrp, mc):
            s = ue_[users_] @ ie_loc.T
            o = torch.sort(s, dim=-1, descending=True, stable=True)
            kk = min(k, s.shape[1])
            vals = torch.full((len(users_), k), float("-inf"))
            idx = torch.full((len(users_), k), np.iinfo(np.int32).max, dtype=torch.int64)
            vals[:, :kk], idx[:, :kk] = o[0][:, :kk], o[1][:, :kk] + item_lo
            return idx, vals
        got_ids = par.sharded_score_topk(ue, users, ie[lo:hi], lo, 20, local_topk=local_topk, merge=_merge_cpu)
        want_ids = torch.sort(ue @ ie.T, dim=-1, descending=True, stable=True)[1][:, :20]
        assert torch.equal(got_ids, want_ids)
        # user-partitioned bipartite propagation (items replicated, one all-reduce per layer)
        def cpu_csr_from_coo(rows, cols, vals, n_rows, n_cols, with_transpose=False):
            o = np.lexsort((cols.numpy(), rows.numpy()))
            rpn = np.concatenate(([0], np.cumsum(np.bincount(rows.numpy(), minlength=n_rows)))).astype(np.int32)
            return pkg("graph").CSRGraph(torch.from_numpy(rpn), cols[o].to(torch.int32), vals[o], n_rows, n_cols)

        def spmm_cpu(g, X_, Y=None, acc_in=None, acc_out=None, scale=1.0):
            rp_ = g.row_ptr.long()
            base = int(rp_[0])
            A_ = torch.sparse_csr_tensor(rp_ - base, g.col_idx.long()[base: int(rp_[-1])] - g.col_offset,
                                         g.vals.double()[base: int(rp_[-1])], size=(g.n_rows, g.n_cols))
            y = torch.sparse.mm(A_, X_.double()[: g.n_cols]).to(X_.dtype)
            if Y is not None:
                Y.copy_(y)
            if acc_out is not None:
                acc_out.copy_(((acc_in + y) if acc_in is not None else y) * scale)

        full.n_users, full.n_items = U, I
        sb = par.ShardedBipartite(full, rank, world, csr_from_coo=cpu_csr_from_coo)
        assert sb.bounds[0] == 0 and sb.bounds[-1] == U
        Xd2 = X.double()
        ou, oi = par.bipartite_propagate_mean(sb, Xd2[sb.lo: sb.hi], Xd2[U:], 3, spmm_fn=spmm_cpu)
        assert torch.allclose(ou, want[sb.lo: sb.hi].detach(), atol=1e-12)
        assert torch.allclose(oi, want[U:].detach(), atol=1e-12)
        # the same shard built from this rank's edges alone (no rank ever holds the full graph):
        # bit-identical values, same result; other chunk counts of the pipelined all-reduce
        eu, ei = torch.from_numpy(np.asarray(u, np.int64)), torch.from_numpy(np.asarray(i, np.int64))
        mine = (eu >= sb.lo) & (eu < sb.hi)
        sb2 = par.ShardedBipartite.from_local_edges(eu[mine], ei[mine], sb.bounds, rank, world, U, I, recipe="f32",
                                                    csr_from_coo=cpu_csr_from_coo)
        assert sb2.nnz_local == sb.nnz_local
        base = int(sb.R.row_ptr[0])
        assert torch.equal(sb2.R.vals, sb.R.vals[base: base + sb.nnz_local])
        for n_chunks in (1, 3, 7):
            ou2, oi2 = par.bipartite_propagate_mean(sb2, Xd2[sb.lo: sb.hi], Xd2[U:], 3, spmm_fn=spmm_cpu, chunks=n_chunks)
            assert torch.allclose(ou2, ou, atol=1e-12) and torch.allclose(oi2, oi, atol=1e-12)
        some_users = torch.tensor([0, U - 1, U // 2, 3, 3])
        got2 = par.bipartite_score_topk(sb, ou, oi, some_users, 10, local_topk=local_topk, merge=_merge_cpu)
        w = want.detach()
        want2 = torch.sort(w[some_users] @ w[U:].T, dim=-1, descending=True, stable=True)[1][:, :10]
        assert torch.equal(got2, want2)
        # item-range sharded feature table: projection forward / backward and the sharded norms
        gen = torch.Generator().manual_seed(11)
        table = torch.randn(I, 24, generator=gen).double()
        Wt, bt = torch.randn(8, 24, generator=gen).double(), torch.randn(8, generator=gen).double()
        Gm = torch.randn(I, 8, generator=gen).double()
        rows = par.ShardedRows(I, rank, world)
        assert rows.lo == min(I, rank * rows.per) and rows.world * rows.per >= I
        xl = rows.local(table).clone().requires_grad_(True)
        Wp, bp = Wt.clone().requires_grad_(True), bt.clone().requires_grad_(True)
        y = par.sharded_projection(xl, Wp, bp, rows, linear_fn=torch.nn.functional.linear)
        tref, Wr, br = (t.clone().requires_grad_(True) for t in (table, Wt, bt))
        yr = torch.nn.functional.linear(tref, Wr, br)
        assert y.shape == yr.shape and torch.allclose(y, yr, atol=1e-12)
        (y * Gm).sum().backward()
        (yr * Gm).sum().backward()
        assert torch.allclose(xl.grad, tref.grad[rows.lo: rows.hi], atol=1e-12)
        assert torch.allclose(Wp.grad, Wr.grad, atol=1e-10) and torch.allclose(bp.grad, br.grad, atol=1e-10)
        ss = par.sharded_sumsq([Wt, bt], [rows.local(table)])
        assert torch.allclose(ss, Wt.pow(2).sum() + bt.pow(2).sum() + table.pow(2).sum(), rtol=1e-12)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_paths_gloo(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_partition_balances_nnz():
    par = pkg("parallel")
    deg = np.concatenate([np.full(1000, 5), np.full(50, 400)])       # users then heavy items
    rp = np.concatenate(([0], np.cumsum(deg)))
    b = par.partition_by_nnz(rp, 4)
    assert b[0] == 0 and b[-1] == 1050 and (np.diff(b) >= 0).all()
    cost = [rp[b[i + 1]] - rp[b[i]] + (b[i + 1] - b[i]) for i in range(4)]
    assert max(cost) < 1.3 * (sum(cost) / 4)
    assert list(par.item_range(10, 3, 4)) == [9, 10] and list(par.item_range(10, 0, 4)) == [0, 3]
