"""One optimiser step of the reference's hot loop (train/train_transformer.py:1104-1298) on the vsn_b200 path:
micro-batch accumulation with `no_sync()` on all but the last micro-batch, soft-target cross entropy,
AdamW (torch fused, as the reference) or SAM(AdamW) two-pass step, EMA update.  Used by bench.py,
`__graft_entry__.smoke()` and the tests; the reference trainer drives the same modules through `dropin/`.
"""
from __future__ import annotations

from contextlib import nullcontext
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .optim import SAM, EMAModel


def soft_target_ce(logits: torch.Tensor, target: torch.Tensor, smoothing: float = 0.1) -> torch.Tensor:
    """regularization/label_smoothing.py:33-77 (reduction='mean'); B x K elements, left to torch."""
    k = logits.shape[-1]
    t = target.to(logits.dtype)
    if smoothing > 0.0:
        t = t * (1.0 - smoothing) + smoothing / k
    return -(t * torch.log_softmax(logits, dim=-1)).sum(-1).mean()


def param_groups(model: torch.nn.Module):
    """utils/helper.py:219-247: no weight decay for biases and 1-D parameters."""
    reg, noreg = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (noreg if name.endswith(".bias") or p.ndim == 1 else reg).append(p)
    return [{"params": reg}, {"params": noreg, "weight_decay": 0.0}]


class TrainStep:
    def __init__(self, model: torch.nn.Module, *, lr: float = 1e-4, weight_decay: float = 0.05, use_sam: bool = False,
                 sam_rho: float = 0.05, use_ema: bool = True, ema_decay: float = 0.999, smoothing: float = 0.1,
                 ddp_model: Optional[torch.nn.Module] = None, grad_sync=None):
        self.module = model                       # the bare module (EMA / parameters)
        self.model = ddp_model if ddp_model is not None else model   # what forward is called on
        self.grad_sync = grad_sync                # vsn_b200.ddp.GradAllReduce (bucketed NCCL all-reduce) or None
        groups = param_groups(model)
        if use_sam:
            self.opt = SAM(groups, torch.optim.AdamW, rho=sam_rho, adaptive=False, lr=lr, weight_decay=weight_decay,
                           fused=True)
        else:
            self.opt = torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, fused=True)
        self.use_sam = use_sam
        self.ema = EMAModel(model=model, decay=ema_decay) if use_ema else None
        self.smoothing = smoothing

    def _accumulate(self, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        n = len(batches)
        total = None
        from .swin import GradAccumulation
        for i, (x, y) in enumerate(batches):
            last = i == n - 1
            GradAccumulation.begin(final=last, first=i == 0)   # block gradients are summed in-kernel over the micro-batches
            if self.grad_sync is not None:
                ctx = nullcontext() if last else self.grad_sync.no_sync()
            else:
                ctx = nullcontext() if last or not hasattr(self.model, "no_sync") else self.model.no_sync()
            with ctx:
                loss = soft_target_ce(self.model(x), y, self.smoothing) / n
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        GradAccumulation.end()
        if self.grad_sync is not None:
            self.grad_sync.finish()               # gradients are now the cross-rank mean
        return total

    def _zero_grad(self) -> None:
        if self.grad_sync is not None:
            self.grad_sync.zero_grad()            # .grad tensors are views into the flat buckets: keep them
        else:
            self.opt.zero_grad(set_to_none=True)

    def step(self, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        """One optimiser step over the given micro-batches; returns the (device) loss of the first pass."""
        loss = self._accumulate(batches)
        if self.use_sam:
            self.opt.first_step(zero_grad=False)
            self._zero_grad()
            self._accumulate(batches)
            self.opt.second_step(zero_grad=False)
            self._zero_grad()
        else:
            self.opt.step()
            self._zero_grad()
        if self.ema is not None:
            self.ema.update(self.module)
        return loss
