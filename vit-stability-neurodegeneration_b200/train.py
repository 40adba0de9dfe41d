"""One optimiser step of the reference's hot loop (train/train_transformer.py:1104-1298) on the vsn_b200 path:
micro-batch accumulation with `no_sync()` on all but the last micro-batch, soft-target cross entropy,
AdamW (torch fused, as the reference) or SAM(AdamW) two-pass step, EMA update.  Used by bench.py,
`__graft_entry__.smoke()` and the tests; the reference trainer drives the same modules through `dropin/`.

Two things the reference's loop cannot do and this one does:

* parameter gradients are accumulated by the kernels straight into one flat arena (`swin.GradSink` +
  `ddp.GradAllReduce`): no per-parameter `grad += new`, no per-block zero fills, one fill per optimiser pass;
* the forward + loss + backward of a micro-batch is captured once in a CUDA graph and replayed (shapes are static;
  DropPath draws its factors from the graph-safe Philox stream), so a micro-batch costs one host call instead of
  ~280 Python -> ctypes launches.  The data-parallel exchange then starts when the last replay has been enqueued.
"""
from __future__ import annotations

from contextlib import nullcontext
from typing import List, Optional, Sequence, Tuple

import torch

from .ddp import GradAllReduce
from .optim import SAM, EMAModel, FusedAdamW
from .swin import GradSink


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
    """`step(batches)` = one optimiser step over the given micro-batches `[(x [B,1,D,H,W], y [B,K] soft labels)]`.

    fuse_micro_batches  run the accumulation micro-batches of a pass as ONE forward/backward over their concatenation.
                The reference accumulates `loss_i / n` over n micro-batches because 8 volumes are what fits its GPUs
                (train/train_transformer.py:1111-1190); with equal micro-batch sizes and per-sample DropPath / LayerNorm
                the summed gradient is the gradient of the mean loss over all n*B volumes -- the same numbers, and on
                180 GB of HBM the 2 x 8 volumes (38 GB of activations) fit at once, so the token dimension of every
                GEMM doubles and every per-launch cost halves.  The data-parallel exchange is unchanged (one per pass).
    fused_adamw optim.FusedAdamW instead of torch.optim.AdamW(fused=True) (identical update)
    ddp_model   a torch DistributedDataParallel wrap of `model`: gradients then travel through autograd and the
                reducer's hooks exactly as in the reference trainer (no in-place accumulation, no graph)
    grad_sync   a `ddp.GradAllReduce` over `model.parameters()` built by the caller (N > 1); without one the step
                builds its own arena (world size 1: nothing is exchanged)
    graph       capture forward + loss + backward of a micro-batch in a CUDA graph and replay it
    scaler      a torch.amp.GradScaler: the loss is scaled before backward and the optimiser steps through it, call for call
                as the reference's fp16 loop does (train/train_transformer.py:1141-1160,1194-1232: `scale(loss).backward()`,
                `scaler.step` / `update`; under SAM `unscale_` -> first_step -> `update` -> second pass ->
                `second_step(scaler=...)`).  Needed in the f16 precision mode (VSN_B200_PRECISION=f16), whose half-precision
                activation gradients underflow without a loss scale; FusedAdamW takes the unscale and the inf check into
                its own pass.
    graph_comm  (N > 1, graph mode) capture a second graph for the LAST micro-batch of a pass with the bucket
                all-reduces inside it: NCCL runs on the communication stream as a forked branch of the graph, each
                bucket as soon as backward has completed it, so the exchange overlaps the rest of backward exactly as
                in the eager loop.  Off (the default): the buckets are reduced after the last replay, exposed.
                Measured on 2 B200 at the round's final state: exposed 18.78 ms per step (+0.41 ms over one GPU), in-graph
                19.13 ms (+0.76 ms) -- every kernel of the backward pass is a persistent grid of one CTA per SM, so an
                NCCL kernel that overlaps them takes SMs away and the grid it displaces waits for it; the "overlap" costs
                more than the 117 MB take on NVLink by themselves (profiles/r2x_*, r2y_*).
    """

    def __init__(self, model: torch.nn.Module, *, lr: float = 1e-4, weight_decay: float = 0.05, use_sam: bool = False,
                 sam_rho: float = 0.05, use_ema: bool = True, ema_decay: float = 0.999, smoothing: float = 0.1,
                 ddp_model: Optional[torch.nn.Module] = None, grad_sync: Optional[GradAllReduce] = None,
                 graph: bool = False, graph_comm: bool = False, fused_adamw: bool = True, fuse_micro_batches: bool = False,
                 scaler: Optional["torch.amp.GradScaler"] = None):
        self.module = model                       # the bare module (EMA / parameters)
        self.model = ddp_model if ddp_model is not None else model   # what forward is called on
        self.autograd_grads = ddp_model is not None
        if self.autograd_grads:
            if graph or grad_sync is not None:
                raise ValueError("torch DDP owns the gradients: no grad_sync / graph with ddp_model")
            self.grad_sync = None
        else:
            self.grad_sync = grad_sync if grad_sync is not None else GradAllReduce(model.parameters())
        groups = param_groups(model)
        # fused_adamw: optim.FusedAdamW (one multi-tensor launch: update + gradient clear); off: torch's fused AdamW, the
        # reference's choice (train/train_transformer.py:2125-2131) -- same update rule, same state layout
        base = FusedAdamW if fused_adamw else torch.optim.AdamW
        if use_sam:
            self.opt = SAM(groups, base, rho=sam_rho, adaptive=False, lr=lr, weight_decay=weight_decay, fused=True)
        else:
            self.opt = base(groups, lr=lr, weight_decay=weight_decay, fused=True)
        self.fused_adamw = fused_adamw
        self.fuse_micro_batches = fuse_micro_batches and ddp_model is None
        self.use_sam = use_sam
        self.scaler = scaler
        self.ema = EMAModel(model=model, decay=ema_decay) if use_ema else None
        self.smoothing = smoothing
        self.use_graph = graph
        self.graph_comm = graph_comm
        self._graph = None            # (torch.cuda.CUDAGraph, static x, static y, static loss)
        self._graph_sync = None       # same micro-batch with the bucket all-reduces captured inside (last micro-batch)
        self._graph_key = None
        self.graph_kernel_nodes = 0
        self.graph_replays = 0
        self._shadows = [m._shadow_owner() for m in model.modules() if hasattr(m, "_shadow_owner")]

    # ------------------------------------------------------------------ one micro-batch
    def _fwd_bwd(self, x: torch.Tensor, y: torch.Tensor, n: int) -> torch.Tensor:
        loss = soft_target_ce(self.model(x), y, self.smoothing) / n
        (loss if self.scaler is None else self.scaler.scale(loss)).backward()
        return loss.detach()

    def _capture(self, x: torch.Tensor, y: torch.Tensor, n: int) -> None:
        """Warm up on a side stream (allocator, plans, lazily built tables), then capture one micro-batch; with
        several ranks and `graph_comm`, capture it a second time with the gradient exchange inside."""
        from . import _lib
        gs = self.grad_sync
        sx, sy = torch.empty_like(x), torch.empty_like(y)
        sx.copy_(x)
        sy.copy_(y)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._fwd_bwd(sx, sy, n)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(g):
            sl = self._fwd_bwd(sx, sy, n)
        self.graph_kernel_nodes = _lib.launch_count() - n0     # library kernels one replay launches
        self._graph = (g, sx, sy, sl)
        self._graph_sync = None
        if gs.world > 1 and self.graph_comm:
            notify = lambda grads: [gs.mark_ready(t) for t in grads]     # noqa: E731
            GradSink.notify = notify
            try:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):                  # eager rehearsal with the collectives (every rank)
                    self._fwd_bwd(sx, sy, n)
                    gs.finish()
                torch.cuda.current_stream().wait_stream(side)
                g2 = torch.cuda.CUDAGraph()
                # the NCCL watchdog thread queries events while we capture: keep the capture thread-local
                with torch.cuda.graph(g2, pool=g.pool(), capture_error_mode="thread_local"):
                    sl2 = self._fwd_bwd(sx, sy, n)
                    gs.finish()                                # joins the communication stream back into the graph
                self._graph_sync = (g2, sx, sy, sl2)
            finally:
                GradSink.notify = None
        gs.zero_grad()                             # warm-up and capture passes left their sums in the arena
        self._graph_key = (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype, n)

    def _refresh_weights(self) -> None:
        for sh in self._shadows:
            sh.refresh(force=True)

    def _fused(self, batches):
        """The micro-batches as one batch (see `fuse_micro_batches`); in graph mode, once the graph exists, they are
        copied straight into their slices of its static input buffers instead of being concatenated first."""
        x0, y0 = batches[0]
        if any(x.shape != x0.shape or x.dtype != x0.dtype or y.shape != y0.shape or y.dtype != y0.dtype
               for x, y in batches[1:]):
            return batches                 # ragged last accumulation group: plain accumulation
        B, n = x0.shape[0], len(batches)
        key = ((n * B,) + tuple(x0.shape[1:]), x0.dtype, (n * B,) + tuple(y0.shape[1:]), y0.dtype, 1)
        if self.use_graph and self._graph is not None and self._graph_key == key:
            sx, sy = self._graph[1], self._graph[2]
            for i, (x, y) in enumerate(batches):
                sx[i * B:(i + 1) * B].copy_(x, non_blocking=True)
                sy[i * B:(i + 1) * B].copy_(y, non_blocking=True)
            return [(sx, sy)]
        return [(torch.cat([x for x, _ in batches]), torch.cat([y for _, y in batches]))]

    def _accumulate(self, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        if self.fuse_micro_batches and len(batches) > 1:
            batches = self._fused(batches)
        n = len(batches)
        if self.autograd_grads:
            total = None
            for i, (x, y) in enumerate(batches):
                ctx = nullcontext() if i == n - 1 or not hasattr(self.model, "no_sync") else self.model.no_sync()
                with ctx:
                    loss = self._fwd_bwd(x, y, n)
                total = loss if total is None else total + loss
            return total
        gs = self.grad_sync
        total = None
        holds = [(sh, sh.hold) for sh in self._shadows]
        GradSink.enabled = True
        try:
            # the bf16 weight shadow is refreshed once per pass: the masters do not change between micro-batches
            self._refresh_weights()
            for sh, _ in holds:
                sh.hold = True
            if self.use_graph:
                GradSink.notify = None
                x0, y0 = batches[0]
                key = (tuple(x0.shape), x0.dtype, tuple(y0.shape), y0.dtype, n)
                for x, y in batches[1:]:
                    if (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype, n) != key:
                        raise ValueError("graph mode needs micro-batches of one shape and dtype")
                if self._graph is None or self._graph_key != key:
                    self._capture(x0, y0, n)       # before anything of this pass has been accumulated
                for i, (x, y) in enumerate(batches):
                    g, sx, sy, sl = self._graph_sync if (i == n - 1 and self._graph_sync is not None) else self._graph
                    if x.data_ptr() != sx.data_ptr():
                        sx.copy_(x, non_blocking=True)
                    if y.data_ptr() != sy.data_ptr():
                        sy.copy_(y, non_blocking=True)
                    g.replay()
                    self.graph_replays += 1
                    total = sl.clone() if total is None else total + sl
                if self._graph_sync is None:
                    gs.reduce_all()               # no collectives in the graph: reduce now (exposed)
            else:
                for i, (x, y) in enumerate(batches):
                    last = i == n - 1
                    # buckets are reduced as the last micro-batch's backward completes them
                    GradSink.notify = (lambda grads: [gs.mark_ready(g) for g in grads]) if last else None
                    loss = self._fwd_bwd(x, y, n)
                    total = loss if total is None else total + loss
        finally:
            GradSink.enabled, GradSink.notify = False, None
            for sh, h in holds:
                sh.hold = h
        gs.finish()                               # gradients are now the cross-rank mean
        return total

    def input_buffers(self) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        """Graph mode: the captured micro-batch's static (x, y) buffers (None before the first step)."""
        return None if self._graph is None else (self._graph[1], self._graph[2])

    def _zero_grad(self) -> None:
        if self.grad_sync is not None:
            self.grad_sync.zero_grad()            # .grad tensors are views into the flat arena: keep them
        else:
            self.opt.zero_grad(set_to_none=True)

    def step(self, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        """One optimiser step over the given micro-batches; returns the (device) loss of the first pass."""
        loss = self._accumulate(batches)
        if self.scaler is not None:
            sc = self.scaler
            if self.use_sam:
                sc.unscale_(self.opt.base_optimizer)
                self.opt.first_step(zero_grad=False)
                self._zero_grad()
                sc.update()
                self._accumulate(batches)
                self.opt.second_step(zero_grad=False, scaler=sc)       # restore, scaler.step(base), scaler.update()
                self._zero_grad()
            else:
                if self.fused_adamw:
                    sc.step(self.opt, zero_grad=True)
                else:
                    sc.step(self.opt)
                    self._zero_grad()
                sc.update()
        elif self.use_sam:
            self.opt.first_step(zero_grad=False)
            self._zero_grad()
            self._accumulate(batches)
            if self.fused_adamw:
                self.opt.second_step(zero_grad=True)       # restore + AdamW; the AdamW pass clears the gradients
            else:
                self.opt.second_step(zero_grad=False)
                self._zero_grad()
        elif self.fused_adamw:
            self.opt.step(zero_grad=True)
        else:
            self.opt.step()
            self._zero_grad()
        if self.ema is not None:
            self.ema.update(self.module)
        return loss
