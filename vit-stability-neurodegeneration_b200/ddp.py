"""Batch-sharded data parallelism: bucketed gradient all-reduce overlapped with backward.

Replaces what `torch.nn.parallel.DistributedDataParallel(...)` does for the reference on this path
(train/train_transformer.py:2099-2108 wraps the model; the reducer then averages the 117 MB of fp32 gradients
over the ranks in ~25 MB buckets while backward is still running, and `model.no_sync()` skips the exchange on
all but the last accumulation micro-batch, train/train_transformer.py:1131-1137,1225-1231).

One process per GPU.  All gradients live in ONE flat fp32 arena cut into buckets; every parameter's `.grad` is a
view into its bucket, so "packing" and "unpacking" cost nothing and `zero_grad()` is a single fill.  Buckets are
laid out in reverse registration order, which is the order backward produces gradients (head + stage 3 first,
patch embedding last); when the last gradient of a bucket is complete a single `all_reduce(AVG)`
of the whole bucket is enqueued on a communication stream (NCCL over NVLink / NVSwitch), overlapping the rest of
backward; `finish()` makes the compute stream wait for the outstanding reductions.  At construction rank 0's
parameters (and the buffers handed in) are broadcast, as DistributedDataParallel does, so replicas that were
initialised differently cannot silently train different weights.  There is no data-path collective anywhere
else: volumes are independent units and the weights are replicated.

Gradients reach the arena in one of two ways: through autograd (`post_accumulate_grad` hooks count a bucket's
parameters down), or written in place by the kernels (`swin.GradSink`, used by `train.TrainStep`), in which
case the producer calls `mark_ready(grad)` itself.

The same host logic runs on CPU tensors over the `gloo` backend (tests/test_ddp_cpu.py, world size 2).
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def plan_buckets(sizes: Sequence[int], cap_elems: int) -> List[List[int]]:
    """Greedy bucket plan over parameter indices in REVERSE order (= gradient production order).
    A parameter larger than the cap gets a bucket of its own.  Returns lists of parameter indices."""
    buckets: List[List[int]] = []
    cur: List[int] = []
    cur_n = 0
    for i in reversed(range(len(sizes))):
        n = int(sizes[i])
        if cur and cur_n + n > cap_elems:
            buckets.append(cur)
            cur, cur_n = [], 0
        cur.append(i)
        cur_n += n
    if cur:
        buckets.append(cur)
    return buckets


class GradAllReduce:
    """Bucketed mean all-reduce of the gradients of `params`, overlapped with backward.

        ddp = GradAllReduce(model.parameters(), buffers=model.buffers())   # after the process group exists
        for micro in micro_batches[:-1]:
            with ddp.no_sync():
                loss(micro).backward()                    # accumulate locally
        loss(micro_batches[-1]).backward()                # buckets are reduced as they fill
        ddp.finish()                                      # grads are now the cross-rank mean
        optimizer.step(); ddp.zero_grad()

    `.grad` of every parameter is a persistent view into the flat arena: use `ddp.zero_grad()` (or
    `optimizer.zero_grad(set_to_none=False)`), never `set_to_none=True`.  With world size 1 (or no process group)
    the object is just that arena: nothing is exchanged.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 25.0,
                 group: Optional[dist.ProcessGroup] = None, buffers: Optional[Iterable[torch.Tensor]] = None,
                 broadcast: bool = True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradAllReduce needs at least one parameter that requires grad")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        dev = self.params[0].device
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise RuntimeError("GradAllReduce expects fp32 parameters on one device")
        self.device = dev
        if self.world > 1 and broadcast:
            # DistributedDataParallel semantics: every replica starts from rank 0's state
            src = dist.get_global_rank(group, 0) if group is not None else 0
            for t in [p.data for p in self.params] + [b for b in (buffers or [])]:
                dist.broadcast(t, src=src, group=group)
        cap = max(1, int(bucket_mb * 1024 * 1024 / 4))
        self.plan = plan_buckets([p.numel() for p in self.params], cap)
        # one arena, buckets are consecutive slices of it; 16-byte aligned parameter slices so vectorised
        # kernels can run on the views
        layout, tot = [], 0
        for idxs in self.plan:
            start, offs = tot, []
            for i in idxs:
                offs.append(tot)
                tot += (self.params[i].numel() + 3) // 4 * 4
            layout.append((start, tot, offs))
        self.arena = torch.zeros(tot, device=dev, dtype=torch.float32)
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._index_of_grad = {}
        self._pending_init: List[int] = []
        for b, (idxs, (start, end, offs)) in enumerate(zip(self.plan, layout)):
            self.buckets.append(self.arena[start:end])
            for i, o in zip(idxs, offs):
                p = self.params[i]
                p.grad = self.arena[o: o + p.numel()].view_as(p)
                self._bucket_of[i] = b
                self._index_of_grad[p.grad.data_ptr()] = i
            self._pending_init.append(len(idxs))
        self._pending = list(self._pending_init)
        self._launched = [False] * len(self.buckets)
        self._sync = True
        self._works = []
        self._comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    # ------------------------------------------------------------------ gradient arrival
    def _make_hook(self, i: int):
        def hook(param: torch.nn.Parameter):
            g = param.grad
            if g is None or g.untyped_storage().data_ptr() != self.arena.untyped_storage().data_ptr():
                raise RuntimeError("a .grad was replaced (zero_grad(set_to_none=True)?): use GradAllReduce.zero_grad()")
            self._arrived(i)
        return hook

    def mark_ready(self, grad: torch.Tensor) -> None:
        """The kernels have finished accumulating into this `.grad` view (enqueued on the current stream)."""
        i = self._index_of_grad.get(grad.data_ptr())
        if i is None:
            raise RuntimeError("mark_ready: not a gradient view of this arena")
        self._arrived(i)

    def _arrived(self, i: int) -> None:
        if not self._sync or self.world == 1:
            return
        b = self._bucket_of[i]
        if self._launched[b]:
            raise RuntimeError(
                "a gradient arrived after its bucket was reduced: two synchronised backward passes without "
                "finish() in between (or a parameter used twice); run the earlier passes under no_sync()")
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int) -> None:
        flat = self.buckets[b]
        self._launched[b] = True
        op = dist.ReduceOp.AVG if self.device.type == "cuda" else dist.ReduceOp.SUM
        if self._comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))       # the bucket's gradients are complete here
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                work = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
            self._works.append(work)
        else:
            self._works.append(dist.all_reduce(flat, op=op, group=self.group, async_op=True))

    # ------------------------------------------------------------------ API
    @contextmanager
    def no_sync(self):
        """Accumulate locally (non-final micro-batches; reference: `model.no_sync()`)."""
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    def reduce_all(self) -> None:
        """Reduce every bucket that has not been reduced yet, in bucket order (the same on every rank).  Used when
        the producer gives no per-gradient notice (a captured CUDA graph) and for parameters that took no part."""
        if self.world > 1:
            for b in range(len(self.buckets)):
                if not self._launched[b]:
                    self._launch(b)

    def finish(self) -> None:
        """Wait for the outstanding reductions: gradients are the cross-rank mean afterwards.  Call after the last
        backward.  Buckets some of whose parameters produced no gradient are reduced here, in fixed bucket order."""
        if self.world > 1:
            self.reduce_all()
            for w in self._works:
                w.wait()            # NCCL: makes the current stream wait for the collective
            if self._comm_stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
            else:                   # gloo has no AVG: sum, then scale
                self.arena.mul_(1.0 / self.world)
        self._works = []
        self._pending = list(self._pending_init)
        self._launched = [False] * len(self.buckets)

    def zero_grad(self) -> None:
        self.arena.zero_()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def bytes_per_reduce(self) -> int:
        return self.arena.numel() * 4
