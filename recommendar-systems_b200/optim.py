"""FusedAdam: torch.optim.Adam semantics (trainer.py:126-143) on one multi-tensor kernel."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def axpy_multi(ys, xs, coef, sign=1.0):
    """y_t += sign * coef * x_t for lists of tensors; `coef` is a 1-element CUDA float tensor."""
    n = len(ys)
    if n:
        lib.call("mmrec_axpy_multi_f32", _ptr_array(ys), _ptr_array(xs),
                 (C.c_int64 * n)(*[t.numel() for t in ys]), n, lib.ptr(coef), float(sign), lib.stream())


def mirror_coef(params, grads, hyper, numel_total, alpha_base, alpha_max_scale, target_rel_step):
    """Device float [2] = (alpha_eff * lr, alpha_eff) of the mirror-gradient step
    (trainer.py:289-305) from one pass over parameters and gradients (mmrec_mirror_coef_f32)."""
    n = len(params)
    numel = (C.c_int64 * n)(*[t.numel() for t in params])
    ws = torch.empty(lib.load().mmrec_mirror_coef_workspace_bytes(numel, n), dtype=torch.uint8,
                     device=params[0].device)
    out = torch.empty(2, dtype=torch.float32, device=params[0].device)
    lib.call("mmrec_mirror_coef_f32", _ptr_array(params), _ptr_array(grads), numel, n, lib.ptr(hyper),
             float(numel_total), float(alpha_base), float(alpha_max_scale), float(target_rel_step), lib.ptr(ws),
             lib.ptr(out), lib.stream())
    return out


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for `optim.Adam(params, lr, weight_decay)` (betas 0.9/0.999, eps 1e-8, no amsgrad).
    State keys `exp_avg` / `exp_avg_sq` as in torch; `param_groups[i]['lr']` is honoured (LambdaLR
    works). The learning rate and the update count live in device memory (`group['hyper']`), so a
    step captured in a CUDA graph replays correctly; consequently the bias-correction step is per
    group, not per parameter (identical to torch whenever every parameter gets a gradient)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _hyper(self, group, device):
        h = group.get("hyper")
        if h is None:
            h = torch.tensor([group["lr"], 0.0], dtype=torch.float64, device=device)
            group["hyper"], group["hyper_lr"] = h, group["lr"]
        return h

    def sync_lr(self):
        """Push a changed `group['lr']` (LR scheduler) to the device. Call outside graph capture."""
        for group in self.param_groups:
            h = group.get("hyper")
            if h is not None and group["hyper_lr"] != group["lr"]:
                h[0:1].fill_(group["lr"])
                group["hyper_lr"] = group["lr"]

    def lr_tensor(self):
        g = self.param_groups[0]
        dev = g["params"][0].device
        return self._hyper(g, dev)[0:1]

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0, undo=None):
        """`grad_scale` multiplies every gradient inside the kernel (fp32, as `_foreach_mul_` would).
        `undo` = ({param: tensor}, coef): every parameter is first moved by coef * tensor (the
        return from the mirror-gradient point) in the same pass."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs, us = [], [], [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                ps.append(p)
                gs.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
                if undo is not None:
                    us.append(undo[0][p])
            n = len(ps)
            if n == 0:
                continue
            if not torch.cuda.is_current_stream_capturing():
                self.sync_lr()
            hyper = self._hyper(group, ps[0].device)
            lib.call("mmrec_adam_step_f32", _ptr_array(ps), _ptr_array(gs), _ptr_array(ms), _ptr_array(vs),
                     (C.c_int64 * n)(*[t.numel() for t in ps]), n, lib.ptr(hyper),
                     float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                     float(group["weight_decay"]), float(grad_scale),
                     _ptr_array(us) if undo is not None else None,
                     lib.ptr(undo[1]) if undo is not None else None, lib.stream())
        return loss
