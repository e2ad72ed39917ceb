"""FusedAdam: torch.optim.Optimizer facade over gmc_adam_multi.

Keeps `torch.optim.Adam`'s state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter;
param_groups with lr/betas/eps/weight_decay/amsgrad/...) so `optimizer.state_dict()` written
into checkpoints (reference python/Training/TrainingNeural.py:447-482) stays loadable by
`torch.optim.Adam.load_state_dict` and vice versa.  Parameters whose `.grad` is None are
skipped exactly like torch does -- that is what keeps the reference's unused nn.Embedding
(:332-336) untouched.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, amsgrad: bool = False):
        if weight_decay != 0.0 or amsgrad:
            raise NotImplementedError("FusedAdam implements the reference's configuration only "
                                      "(weight_decay=0, amsgrad=False)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=False)
        super().__init__(params, defaults)

    def _state_for(self, p: torch.Tensor) -> dict:
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def fused_step(self, params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], shadows=None) -> None:
        """One Adam step on `params` with explicit gradient tensors (engine-owned buffers).  shadows: optional list of
        contiguous bf16 tensors (or None) that receive bf16 copies of the updated parameters in the same pass."""
        by_group = {}
        for group in self.param_groups:
            for p in group["params"]:
                by_group[id(p)] = group
        buckets = {}
        for i, (p, g) in enumerate(zip(params, grads)):
            group = by_group[id(p)]
            st = self._state_for(p)
            st["step"] += 1
            key = (id(group), int(st["step"].item()))
            buckets.setdefault(key, (group, [], [], [], [], []))
            _, ps, gs, ms, vs, shs = buckets[key]
            ps.append(p.data); gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            shs.append(shadows[i] if shadows is not None else None)
        for (_, step), (group, ps, gs, ms, vs, shs) in buckets.items():
            b1, b2 = group["betas"]
            ops.adam_multi(ps, gs, ms, vs, lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"], step=step,
                           shadows=shs if any(sh is not None for sh in shs) else None)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        params: List[torch.Tensor] = []
        grads: List[torch.Tensor] = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam: parameter with a gradient lives on the CPU; the optimiser is CUDA-only")
                params.append(p)
                grads.append(p.grad.contiguous())
        if params:
            self.fused_step(params, grads)
        return loss
