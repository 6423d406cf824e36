"""Device-resident CSR graphs for the propagation kernels.

HBM layout: `row_ptr` int32 [rows+1], `col_idx` int32 [nnz], `vals` float32
[nnz], plus the SpMM work list `tasks` int32 [n_tasks, 4] (<= 64 non-zeros per task, see
build_tasks) with its reduction slots for heavy rows. The symmetric user-item adjacency is stored once; R = A[:U, U:] and R^T are
zero-copy views of it (a row-pointer slice + a column offset), so SMORE/MGCN's `R` propagation
and its transposed backward need no extra storage.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib

SEG = 64            # non-zeros per SpMM task (kSeg in csrc/spmm.cu)


def build_tasks(row_ptr, skip_empty=False):
    """Work list of mmrec_spmm_csr_f32: int32 [n_tasks, 4] = {row, begin, end, slot}. Rows with at
    most SEG non-zeros are one task (slot -1, longest first so that the sub-warps of a warp have
    similar trip counts); heavier rows are cut into SEG-sized parts that share a reduction slot.
    `skip_empty` leaves rows without non-zeros out (their output row is then not written at all:
    only for accumulating launches, see ColumnBlockedCSR).
    Returns (tasks, slot_base int32 [n_heavy], total_parts)."""
    dev = row_ptr.device
    rp = row_ptr.to(torch.int64)
    begin, end = rp[:-1], rp[1:]
    deg = end - begin
    rows = torch.arange(deg.numel(), device=dev, dtype=torch.int64)
    heavy = deg > SEG
    light_rows = rows[~heavy & ((deg > 0) if skip_empty else (deg >= 0))]
    order = torch.argsort(deg[light_rows], descending=True, stable=True)
    light_rows = light_rows[order]
    light = torch.stack([light_rows, begin[light_rows], end[light_rows],
                         torch.full_like(light_rows, -1)], dim=1)
    h_rows = rows[heavy]
    n_parts = (deg[h_rows] + SEG - 1) // SEG
    slot_base = torch.cumsum(n_parts, 0) - n_parts
    total_parts = int(n_parts.sum().item()) if h_rows.numel() else 0
    if total_parts:
        slot = torch.repeat_interleave(torch.arange(h_rows.numel(), device=dev), n_parts)
        part = torch.arange(total_parts, device=dev) - slot_base[slot]
        r = h_rows[slot]
        b = begin[r] + part * SEG
        e = torch.minimum(b + SEG, end[r])
        heavy_tasks = torch.stack([r, b, e, slot], dim=1)
        tasks = torch.cat([heavy_tasks, light], dim=0)
    else:
        tasks = light
    return tasks.to(torch.int32).contiguous(), slot_base.to(torch.int32).contiguous(), total_parts


class CSRGraph:
    def __init__(self, row_ptr, col_idx, vals, n_rows, n_cols, col_offset=0, symmetric=False, skip_empty=False):
        self.row_ptr, self.col_idx, self.vals = row_ptr, col_idx, vals
        self.n_rows, self.n_cols, self.col_offset = int(n_rows), int(n_cols), int(col_offset)
        self.tasks, self.slot_base, self.total_parts = build_tasks(row_ptr, skip_empty)
        self.n_tasks = int(self.tasks.shape[0])
        self.counters = torch.zeros(max(1, self.slot_base.numel()), dtype=torch.int32,
                                    device=row_ptr.device)
        self._scratch = {}
        self.nnz = int((row_ptr[-1] - row_ptr[0]).item())
        self.t = self if symmetric else None

    def scratch(self, d, slot=0):
        """Partial-sum buffer of the heavy rows for embedding width d (allocated once). `slot`
        separates concurrent uses of the same graph inside one multi-problem launch."""
        if (d, slot) not in self._scratch:
            self._scratch[(d, slot)] = torch.empty(max(1, self.total_parts) * d, dtype=torch.float32,
                                                   device=self.vals.device)
        return self._scratch[(d, slot)]

    def counters_for(self, slot=0):
        """Self-resetting arrival counters of the heavy rows; one set per concurrent use."""
        if slot == 0:
            return self.counters
        extra = self.__dict__.setdefault("_extra_counters", {})
        if slot not in extra:
            extra[slot] = torch.zeros_like(self.counters)
        return extra[slot]

    @property
    def device(self):
        return self.vals.device

    def algorithmic_bytes(self, d):
        """8*nnz + 4*(rows+1) + 4*d*(rows + cols)."""
        return 8 * self.nnz + 4 * (self.n_rows + 1) + 4 * d * (self.n_rows + self.n_cols)

    def to_torch_coo(self):
        """For tests: (rows int64, cols int64, vals) on the host."""
        rp = self.row_ptr.cpu().numpy().astype(np.int64)
        base = rp[0]
        counts = np.diff(rp)
        rows = np.repeat(np.arange(self.n_rows, dtype=np.int64), counts)
        cols = self.col_idx.cpu().numpy()[base:base + counts.sum()].astype(np.int64) - self.col_offset
        vals = self.vals.cpu().numpy()[base:base + counts.sum()]
        return rows, cols, vals


def _degree_lut(recipe, n):
    """deg -> deg^-1/2 with the reference's own host arithmetic (bit-exact by construction)."""
    if recipe == "f64eps":        # layergcn.py:103-107: np.power(deg + 1e-7, -0.5) in float64
        return np.power(np.arange(n, dtype=np.float64) + 1e-7, -0.5), True
    if recipe == "f32":           # mgcn.py:120-123: np.power(rowsum_f32, -0.5), inf -> 0
        with np.errstate(divide="ignore"):
            t = np.power(np.arange(n, dtype=np.float32), np.float32(-0.5)).astype(np.float32)
        t[np.isinf(t)] = 0.0
        return t, False
    if recipe == "edge_f32":      # layergcn.py:72-81: torch.pow(1e-7 + deg_int64, -0.5) float32
        return torch.pow(1e-7 + torch.arange(n, dtype=torch.int64), -0.5).numpy(), False
    raise ValueError(f"unknown normalisation recipe {recipe!r}")


def build_ui_graph(users, items, n_users, n_items, recipe):
    """K11/K12: CSR of D^-1/2 [[0,R],[R^T,0]] D^-1/2 from unique (user, item) edges on device."""
    lib.require_cuda(users, items)
    dev = users.device
    users = users.to(torch.int64).contiguous()
    items = items.to(torch.int64).contiguous()
    E = users.numel()
    n = n_users + n_items
    lut_np, is64 = _degree_lut(recipe, max(n_users, n_items) + 1)
    lut = torch.from_numpy(lut_np).to(dev)
    row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    col_idx = torch.empty(2 * E, dtype=torch.int32, device=dev)
    vals = torch.empty(2 * E, dtype=torch.float32, device=dev)
    L = lib.load()
    ws_bytes = L.mmrec_ui_adj_workspace_bytes(E, n_users, n_items)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    lib.call("mmrec_ui_adj_build", lib.ptr(users), lib.ptr(items), E, n_users, n_items, lib.ptr(lut),
             lut.numel(), int(is64), lib.ptr(row_ptr), lib.ptr(col_idx), lib.ptr(vals), None,
             lib.ptr(ws), ws_bytes, lib.stream())
    g = CSRGraph(row_ptr, col_idx, vals, n, n, symmetric=True)
    g.n_users, g.n_items = n_users, n_items
    return g


def ui_blocks(g):
    """(R, R^T) views of a user-item graph: R = A[:U, U:] (mgcn.py:134, smore.py:198)."""
    U, I = g.n_users, g.n_items
    R = CSRGraph(g.row_ptr[:U + 1], g.col_idx, g.vals, U, I, col_offset=U)
    Rt = CSRGraph(g.row_ptr[U:], g.col_idx, g.vals, I, U, col_offset=0)
    R.t, Rt.t = Rt, R
    return R, Rt


def csr_from_coo(rows, cols, vals, n_rows, n_cols, with_transpose=True, skip_empty=False):
    """COO (int64, unsorted, duplicates kept) -> CSRGraph; optionally with its transpose for
    the backward of non-symmetric graphs."""
    lib.require_cuda(rows, cols, vals)
    dev = rows.device
    rows = rows.to(torch.int64).contiguous()
    cols = cols.to(torch.int64).contiguous()
    vals = vals.to(torch.float32).contiguous()
    nnz = rows.numel()
    L = lib.load()
    ws_bytes = L.mmrec_csr_from_coo_workspace_bytes(nnz)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def one(transpose):
        out_rows = n_cols if transpose else n_rows
        out_cols = n_rows if transpose else n_cols
        row_ptr = torch.empty(out_rows + 1, dtype=torch.int32, device=dev)
        col_idx = torch.empty(nnz, dtype=torch.int32, device=dev)
        out_vals = torch.empty(nnz, dtype=torch.float32, device=dev)
        lib.call("mmrec_csr_from_coo", lib.ptr(rows), lib.ptr(cols), lib.ptr(vals), nnz, n_rows,
                 n_cols, int(transpose), lib.ptr(row_ptr), lib.ptr(col_idx), lib.ptr(out_vals), None,
                 lib.ptr(ws), ws_bytes, lib.stream())
        return CSRGraph(row_ptr, col_idx, out_vals, out_rows, out_cols, skip_empty=skip_empty)

    g = one(False)
    if with_transpose:
        g.t = one(True)
        g.t.t = g
    return g


def from_torch_sparse(a, with_transpose=True):
    """A torch sparse COO tensor (as the reference builds them) -> CSRGraph."""
    idx, val = a._indices(), a._values()
    return csr_from_coo(idx[0], idx[1], val, a.shape[0], a.shape[1], with_transpose)


class ColumnBlockedCSR:
    """A [n_rows, n_cols] sparse matrix cut into column ranges, one CSRGraph per range, for
    operands X that do not fit the L2 cache (126 MB on B200).

    A row-split SpMM gathers one X row per non-zero. When X is larger than L2 and the column
    pattern has no locality (user vectors seen from the item side of a recommendation graph: every
    user row is wanted by ~deg different item rows, far apart in time) every gather is a DRAM
    access: nnz * 4 d bytes instead of the 4 d * n_cols the operand actually holds. Cut into column
    blocks whose X slice stays L2-resident, block b costs its X slice once plus a read-modify-write
    of the output rows it touches (Y = A_0 X_0, then Y += A_b X_b through the accumulate epilogue
    of the same kernel): DRAM traffic falls from nnz * 4 d to ~(n_cols + 2 * #(row, block) pairs) * 4 d.
    Every block keeps GLOBAL column ids, so X is passed whole. Block 0 lists every row (it
    initialises Y); later blocks list only the rows they touch."""

    def __init__(self, blocks, n_rows, n_cols):
        self.blocks, self.n_rows, self.n_cols = list(blocks), int(n_rows), int(n_cols)
        self.nnz = sum(g.nnz for g in self.blocks)

    @classmethod
    def from_col_sorted_coo(cls, rows, cols, vals, n_rows, n_cols, block_cols, build=None):
        """`cols` ascending (e.g. the edge list of a user-sorted interaction graph seen from the item
        side): a column block is a contiguous slice of the three arrays."""
        build = build or csr_from_coo
        edges = torch.arange(0, n_cols + block_cols, block_cols, device=cols.device, dtype=cols.dtype)
        cut = torch.searchsorted(cols.contiguous(), edges).tolist()
        blocks = []
        for b in range(len(cut) - 1):
            lo, hi = cut[b], cut[b + 1]
            if hi == lo and b > 0:
                continue
            blocks.append(build(rows[lo:hi], cols[lo:hi], vals[lo:hi], n_rows, n_cols, with_transpose=False,
                                skip_empty=b > 0))
        return cls(blocks, n_rows, n_cols)

    def algorithmic_bytes(self, d):
        return 8 * self.nnz + 4 * (self.n_rows + 1) + 4 * d * (self.n_rows + self.n_cols)
