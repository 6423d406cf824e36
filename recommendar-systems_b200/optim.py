"""FusedAdam: torch.optim.Adam semantics (trainer.py:126-143) on one multi-tensor kernel."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib


_ptr_array = lib.ptr_array


def axpy_multi(ys, xs, coef, sign=1.0):
    """y_t += sign * coef * x_t for lists of tensors; `coef` is a 1-element CUDA float tensor."""
    n = len(ys)
    if n:
        lib.call("mmrec_axpy_multi_f32", _ptr_array(ys), _ptr_array(xs),
                 (C.c_int64 * n)(*[t.numel() for t in ys]), n, lib.ptr(coef), float(sign), lib.stream())


def mirror_coef(params, grads, hyper, numel_total, alpha_base, alpha_max_scale, target_rel_step, extra=None,
                n_extra_p2=0, n_extra_g2=0):
    """Device float [2] = (alpha_eff * lr, alpha_eff) of the mirror-gradient step
    (trainer.py:289-305) from one pass over parameters and gradients (mmrec_mirror_coef_f32).
    `extra` (device float64): sums of squares of tensors that are not in the lists -- the feature
    tables with a low-rank gradient -- parameters first, then gradients."""
    n = len(params)
    numel = (C.c_int64 * n)(*[t.numel() for t in params])
    ws = torch.empty(lib.load().mmrec_mirror_coef_workspace_bytes(numel, n), dtype=torch.uint8,
                     device=params[0].device)
    out = torch.empty(2, dtype=torch.float32, device=params[0].device)
    lib.call("mmrec_mirror_coef_f32", _ptr_array(params), _ptr_array(grads), numel, n, lib.ptr(hyper),
             float(numel_total), float(alpha_base), float(alpha_max_scale), float(target_rel_step),
             lib.ptr(extra), int(n_extra_p2), int(n_extra_g2), lib.ptr(ws), lib.ptr(out), lib.stream())
    return out


def _lowrank_ws(p):
    n = lib.load().mmrec_table_lowrank_workspace_bytes(p.shape[0], p.shape[1])
    return torch.empty(n, dtype=torch.uint8, device=p.device)


def lowrank_sumsq(p, out):
    """out[0] (device float64) = ||dY W||_F^2 of the table's pending low-rank gradient."""
    lr = p._mmrec_lowrank
    lib.call("mmrec_table_lowrank_sumsq_f64", lib.ptr(lr.dY), lib.ptr(lr.W), p.shape[0], p.shape[1],
             lr.dY.shape[1], lib.ptr(_lowrank_ws(p)), lib.ptr(out), lib.stream())


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

    def zero_grad(self, set_to_none=True):
        """Also drops pending low-rank table gradients (ops.LowRankGrad on `p._mmrec_lowrank`)."""
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "_mmrec_lowrank", None) is not None:
                    p._mmrec_lowrank = None
        return super().zero_grad(set_to_none=set_to_none)

    def lowrank_params(self):
        return [p for g in self.param_groups for p in g["params"] if getattr(p, "_mmrec_lowrank", None) is not None]

    def state_dict(self):
        """torch.optim.Adam layout: every per-parameter state gets the `step` torch keeps there (here
        the update count is per group, on the device, so that captured graphs replay correctly)."""
        sd = super().state_dict()
        for group, packed in zip(self.param_groups, sd["param_groups"]):
            h = group.get("hyper")
            step = float(h[1].item()) if h is not None else 0.0
            for idx in packed["params"]:
                if idx in sd["state"]:
                    sd["state"][idx] = dict(sd["state"][idx], step=torch.tensor(step))
        return sd

    def load_state_dict(self, state_dict):
        """Accepts torch.optim.Adam state: the bias-correction count restarts from the saved `step`
        (the largest one of a group; torch counts per parameter, identical whenever every parameter
        gets a gradient in every step) instead of silently from zero."""
        super().load_state_dict(state_dict)
        for st in self.state.values():
            st.pop("sumsq", None)                    # scratch (float64), re-created by the next step
        for group in self.param_groups:
            steps = [float(self.state[p].pop("step")) for p in group["params"]
                     if p in self.state and "step" in self.state[p]]
            # the device tensor itself is kept (captured graphs and the in-kernel dropout streams hold
            # its address); only its contents follow the loaded state
            h = group.get("hyper")
            if h is not None:
                h[0:1].fill_(group["lr"])
                h[1:2].fill_(max(steps) if steps else 0.0)
                group["hyper_lr"] = group["lr"]
            elif steps and group["params"]:
                self._hyper(group, group["params"][0].device)[1:2].fill_(max(steps))

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0, undo=None):
        """`grad_scale` multiplies every gradient inside the kernel (fp32, as `_foreach_mul_` would).
        `undo` = ({param: tensor}, coef): every parameter is first moved by coef * tensor (the
        return from the mirror-gradient point) in the same pass. Feature tables that carry a
        low-rank gradient (`p._mmrec_lowrank`) are updated first, by the tcgen05 kernel that
        rebuilds each gradient tile from its factors (the factor W is itself a parameter and is
        updated by the multi-tensor launch that follows)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs, us = [], [], [], [], []
            ticked = False
            for p in group["params"]:
                lr = getattr(p, "_mmrec_lowrank", None)
                if lr is None:
                    continue
                if p.grad is not None:
                    raise RuntimeError("FusedAdam: a parameter has both a dense and a low-rank gradient")
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if "sumsq" not in st:
                    st["sumsq"] = torch.zeros(1, dtype=torch.float64, device=p.device)
                if not ticked and not torch.cuda.is_current_stream_capturing():
                    self.sync_lr()
                hyper = self._hyper(group, p.device)
                lib.call("mmrec_table_adam_lowrank_f32", lib.ptr(p), lib.ptr(st["exp_avg"]), lib.ptr(st["exp_avg_sq"]),
                         lib.ptr(lr.dY), lib.ptr(lr.W), p.shape[0], p.shape[1], lr.dY.shape[1], lib.ptr(hyper),
                         float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                         float(group["weight_decay"]), float(grad_scale), int(not ticked), lib.ptr(st["sumsq"]),
                         lib.ptr(_lowrank_ws(p)), lib.stream())
                ticked = True
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
            if not ticked and not torch.cuda.is_current_stream_capturing():
                self.sync_lr()
            hyper = self._hyper(group, ps[0].device)
            lib.call("mmrec_adam_step_f32", _ptr_array(ps), _ptr_array(gs), _ptr_array(ms), _ptr_array(vs),
                     (C.c_int64 * n)(*[t.numel() for t in ps]), n, lib.ptr(hyper),
                     float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                     float(group["weight_decay"]), float(grad_scale),
                     _ptr_array(us) if undo is not None else None,
                     lib.ptr(undo[1]) if undo is not None else None, int(not ticked), lib.stream())
        return loss
