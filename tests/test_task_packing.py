"""CPU checks of graph.build_tasks(pack=True): the packed SpMM work list (MMREC_SPMM_PACKED) covers every
row exactly once, packed tasks are runs of consecutive short rows within the per-task limits, heavy rows keep
their SEG-sized parts."""
import numpy as np
import torch

from conftest import pkg


def test_packed_task_list_invariants():
    G = pkg("graph")
    rng = np.random.default_rng(0)
    for trial in range(30):
        n = int(rng.integers(1, 500))
        deg = rng.choice([0, 0, 1, 2, 3, 5, 8, 13, 20, 31, 32, 33, 40, 64, 65, 200], size=n)
        base = int(rng.integers(0, 50))                      # a row_ptr slice that does not start at 0 (R^T views)
        rp = torch.tensor(np.concatenate(([0], np.cumsum(deg))) + base, dtype=torch.int32)
        tasks, slot_base, total = G.build_tasks(rp, pack=True)
        t = tasks.numpy()
        cover = np.zeros(n, int)
        for row, b, e, w in t:
            if w >= 0:
                continue
            nr = -w
            assert 1 <= nr <= G.PACK_ROWS and e - b <= G.SEG
            assert b == rp[row] and e == rp[row + nr]
            if nr > 1:
                assert (deg[row:row + nr] <= G.PACK_MAX_DEG).all()
            cover[row:row + nr] += 1
        heavy = deg > G.SEG
        assert (cover[~heavy] == 1).all() and (cover[heavy] == 0).all()
        for row in np.flatnonzero(heavy):
            parts = t[(t[:, 0] == row) & (t[:, 3] >= 0)]
            assert len(parts) == -(-deg[row] // G.SEG) and (np.diff(parts[:, 1]) == G.SEG).all()
            assert parts[0, 1] == rp[row] and parts[-1, 2] == rp[row + 1]
        assert total == int(sum(-(-d // G.SEG) for d in deg[heavy]))
        # light tasks are ordered longest first
        light = t[t[:, 3] < 0]
        assert (np.diff(light[:, 2] - light[:, 1]) <= 0).all()
        # the one-row list is what pack=False gives
        t0, _, _ = G.build_tasks(rp, pack=False)
        assert (t0[:, 3].numpy() >= -1).all() and int((t0[:, 3] == -1).sum()) == int((~heavy).sum())


def test_packing_shrinks_sparse_graph_lists():
    G = pkg("graph")
    rng = np.random.default_rng(1)
    deg = rng.poisson(12, size=26495)                        # Baby-like: 12 non-zeros per row
    rp = torch.tensor(np.concatenate(([0], np.cumsum(deg))), dtype=torch.int32)
    a, _, _ = G.build_tasks(rp, pack=True)
    b, _, _ = G.build_tasks(rp, pack=False)
    assert b.shape[0] == 26495 and a.shape[0] < 0.45 * b.shape[0]
    empty = torch.zeros(101, dtype=torch.int32)             # a graph of empty rows: ceil(100 / PACK_ROWS) tasks
    e, _, _ = G.build_tasks(empty, pack=True)
    assert e.shape[0] == -(-100 // G.PACK_ROWS) and int((-e[:, 3]).sum()) == 100
