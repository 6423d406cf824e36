"""Oracle (test infrastructure): the reference's training step restated on torch CPU.

Follows Trainer._train_epoch (/root/reference/src/common/trainer.py:186-335): zero_grad,
calculate_loss, backward, Adam step, and -- for models with `mg_enable` -- the mirror-gradient
block that runs whenever `global_step % mg_interval == 0` (two more forward/backward passes, a
perturbation theta - alpha_eff*lr*g, gradients scaled by -beta, a second Adam step). Because
calculate_loss itself bumps global_step (smore.py:393) every step after the second is a
mirror-gradient step; the same counter logic is kept here.

Used by tests (trajectory parity with the golden `fit/*` vectors) and by bench.py as the timed CPU
baseline (`cpu_baseline` / `--impl reference`).
"""
from __future__ import annotations

import torch

from . import models


class OracleTrainer:
    def __init__(self, model_name, params, graphs, cfg, lr=1e-3, lr_scheduler=(1.0, 50),
                 mg_enable=None, mg_interval=3, mg_alpha=0.5, mg_beta=0.2, mg_target_rel_step=1e-3,
                 mg_alpha_max_scale=20.0, dropout=None):
        self.name, self.G, self.cfg = model_name, graphs, cfg
        self.P = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        self.opt = torch.optim.Adam(list(self.P.values()), lr=lr, weight_decay=0.0)
        self.sched = torch.optim.lr_scheduler.LambdaLR(
            self.opt, lr_lambda=lambda e: lr_scheduler[0] ** (e / lr_scheduler[1]))
        self.mg_enable = (model_name == "SMORE") if mg_enable is None else mg_enable
        self.mg_interval, self.mg_alpha, self.mg_beta = mg_interval, mg_alpha, mg_beta
        self.mg_target_rel_step, self.mg_alpha_max_scale = mg_target_rel_step, mg_alpha_max_scale
        self.global_step = 0
        self.dropout = dropout

    def calculate_loss(self, batch):
        if self.name == "SMORE":
            loss = models.smore_loss(self.P, self.G, self.cfg, batch, dropout=self.dropout)
            self.global_step += 1
            return loss
        return models.LOSS[self.name](self.P, self.G, self.cfg, batch)

    def step(self, batch):
        """One iteration of the batch loop; returns the (first) loss as a float."""
        opt = self.opt
        opt.zero_grad(set_to_none=True)
        loss = self.calculate_loss(batch)
        value = loss.item()
        loss.backward()
        opt.step()
        if self.mg_enable and self.mg_interval > 0 and self.global_step % self.mg_interval == 0:
            lr = opt.param_groups[0]["lr"]
            opt.zero_grad(set_to_none=True)
            self.calculate_loss(batch).backward()
            params = [p for p in self.P.values() if p.grad is not None]
            grads = [p.grad.detach().clone() for p in params]
            with torch.no_grad():
                g_all = torch.cat([g.view(-1) for g in grads])
                grad_rms = float(g_all.norm() / (g_all.numel() ** 0.5))
                p_all = torch.cat([p.detach().view(-1) for p in params])
                param_rms = float(p_all.norm() / (p_all.numel() ** 0.5) + 1e-12)
                alpha = max(self.mg_alpha, self.mg_target_rel_step * param_rms / (lr * grad_rms + 1e-12))
                alpha = min(alpha, self.mg_alpha * self.mg_alpha_max_scale)
                for p, g in zip(params, grads):
                    p.add_(-alpha * lr * g)
            opt.zero_grad(set_to_none=True)
            self.calculate_loss(batch).backward()
            with torch.no_grad():
                for p in self.P.values():
                    if p.grad is not None:
                        p.grad.mul_(-self.mg_beta)
                for p, g in zip(params, grads):
                    p.add_(alpha * lr * g)
            opt.step()
            opt.zero_grad(set_to_none=True)
        return value

    def train_epoch(self, loader):
        total = 0.0
        for batch in loader:
            total += self.step(batch)
        self.sched.step()
        return total

    @torch.no_grad()
    def embeddings(self):
        return models.FORWARD[self.name](self.P, self.G, self.cfg)
