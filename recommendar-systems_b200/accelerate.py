"""`accelerate()`: run UNMODIFIED reference model files on the B200 propagation kernel.

SURVEY.md section 8(b): the reference has no operator layer, its arithmetic sits behind torch call
sites. Every propagation layer of LayerGCN / FREEDOM / MGCN / SMORE / LightGCN is
`torch.sparse.mm(adj, x)` (layergcn.py:133, freedom.py:169,174, mgcn.py:162-184,
smore.py:282-317, lightgcn.py:122) on a CUDA float32 COO adjacency that is uncoalesced, carries
int64 indices and is re-sorted + converted to CSR by ATen on every call, forward and backward.
Inside this context those calls are rerouted to `ops.spmm` (mmrec_spmm_csr_f32): the adjacency is
converted ONCE to our int32 CSR (+ transpose for the backward) and cached per tensor, the backward
runs through our autograd Function. Everything else keeps running on torch.

    with accelerate():
        loss = reference_model.calculate_loss(batch)     # unmodified src/models/smore.py
        loss.backward()

Only calls the kernel covers are taken (sparse COO/CSR float32 CUDA matrix without grad x dense
float32 CUDA [n, d], d in {32, 64, 128, 256}); anything else falls through to torch unchanged --
this is an interception layer, not a fallback: entering the context raises if the CUDA library is
missing.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch.overrides import TorchFunctionMode

from . import graph as G
from . import lib, ops

_WIDTHS = (32, 64, 128, 256)


def _eligible(a, x):
    if not (isinstance(a, torch.Tensor) and isinstance(x, torch.Tensor)):
        return False
    if a.layout not in (torch.sparse_coo, torch.sparse_csr) or x.layout != torch.strided:
        return False
    return (a.is_cuda and x.is_cuda and a.dtype == torch.float32 and x.dtype == torch.float32
            and not a.requires_grad and a.dim() == 2 and x.dim() == 2 and x.shape[1] in _WIDTHS
            and a.shape[1] == x.shape[0])


class accelerate(TorchFunctionMode):
    """TorchFunctionMode that sends `torch.sparse.mm` / `torch.mm` / `torch.matmul` / `Tensor.mm`
    with a sparse left operand to the CSR SpMM kernel. `stats` counts taken and passed-through
    calls and CSR conversions."""

    _FUNCS = None

    def __init__(self, max_cached=32):
        super().__init__()
        self.max_cached = int(max_cached)
        self._cache = OrderedDict()          # key -> (CSRGraph, tensors kept alive)
        self.stats = {"spmm": 0, "converted": 0, "passed": 0}
        if accelerate._FUNCS is None:
            accelerate._FUNCS = {torch.sparse.mm, torch.mm, torch.matmul, torch.Tensor.mm,
                                 torch.Tensor.matmul, torch.spmm}

    def __enter__(self):
        lib.load()                           # no library, no acceleration: fail here, loudly
        return super().__enter__()

    def _graph(self, a):
        """CSRGraph of `a`, converted once. The key pins the identity AND the version of the
        index/value tensors; the entry keeps them alive so their addresses cannot be recycled
        for a different matrix (LayerGCN / FREEDOM build a new adjacency every epoch)."""
        if a.layout == torch.sparse_csr:
            parts = (a.crow_indices(), a.col_indices(), a.values())
        else:
            parts = (a._indices(), a._values())
        key = (a.layout, tuple(a.shape)) + tuple((p.data_ptr(), p._version, p.numel()) for p in parts)
        hit = self._cache.get(key)
        if hit is not None:
            self._cache.move_to_end(key)
            return hit[0]
        if a.layout == torch.sparse_csr:
            crow, col, val = parts
            rows = torch.repeat_interleave(torch.arange(a.shape[0], device=a.device), crow[1:] - crow[:-1])
            g = G.csr_from_coo(rows, col, val, a.shape[0], a.shape[1], with_transpose=True)
        else:
            g = G.from_torch_sparse(a, with_transpose=True)
        self.stats["converted"] += 1
        self._cache[key] = (g, parts)
        while len(self._cache) > self.max_cached:
            self._cache.popitem(last=False)
        return g

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in accelerate._FUNCS and len(args) == 2 and not kwargs and _eligible(args[0], args[1]):
            g = self._graph(args[0])
            self.stats["spmm"] += 1
            return ops.spmm(g, args[1])
        if func in (torch.sparse.mm, torch.spmm):
            self.stats["passed"] += 1
        return func(*args, **kwargs)
