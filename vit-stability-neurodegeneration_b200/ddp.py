"""Batch-sharded data parallelism: bucketed gradient all-reduce overlapped with backward.

Replaces what `torch.nn.parallel.DistributedDataParallel(...)` does for the reference on this path
(train/train_transformer.py:2099-2108 wraps the model; the reducer then averages the 117 MB of fp32 gradients
over the ranks in ~25 MB buckets while backward is still running, and `model.no_sync()` skips the exchange on
all but the last accumulation micro-batch, train/train_transformer.py:1131-1137,1225-1231).

One process per GPU.  Gradients live in a few flat fp32 buckets; every parameter's `.grad` is a view into its
bucket, so "packing" and "unpacking" cost nothing.  Buckets are laid out in reverse registration order, which
is the order backward produces gradients (head + stage 3 first, patch embedding last); when the
last gradient of a bucket has been accumulated a single `all_reduce(SUM)` of the whole bucket is enqueued on a
communication stream (NCCL over NVLink / NVSwitch), overlapping the rest of backward; `finish()` makes the
compute stream wait for the outstanding reductions and applies the 1/world mean.  There is no data-path
collective anywhere else: volumes are independent units and the weights are replicated.

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

        ddp = GradAllReduce(model.parameters())          # after the process group exists
        for micro in micro_batches[:-1]:
            with ddp.no_sync():
                loss(micro).backward()                    # accumulate locally
        loss(micro_batches[-1]).backward()                # buckets are reduced as they fill
        ddp.finish()                                      # grads are now the cross-rank mean
        optimizer.step(); ddp.zero_grad()

    `.grad` of every parameter is a persistent view into a flat bucket: use `ddp.zero_grad()` (or
    `optimizer.zero_grad(set_to_none=False)`), never `set_to_none=True`.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 25.0,
                 group: Optional[dist.ProcessGroup] = None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradAllReduce needs at least one parameter that requires grad")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = self.params[0].device
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise RuntimeError("GradAllReduce expects fp32 parameters on one device")
        self.device = dev
        cap = max(1, int(bucket_mb * 1024 * 1024 / 4))
        self.plan = plan_buckets([p.numel() for p in self.params], cap)
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._pending_init: List[int] = []
        for b, idxs in enumerate(self.plan):
            # 16-byte aligned slices so vectorised kernels can run on the views
            offs, tot = [], 0
            for i in idxs:
                offs.append(tot)
                tot += (self.params[i].numel() + 3) // 4 * 4
            flat = torch.zeros(tot, device=dev, dtype=torch.float32)
            self.buckets.append(flat)
            for i, o in zip(idxs, offs):
                p = self.params[i]
                p.grad = flat[o: o + p.numel()].view_as(p)
                self._bucket_of[i] = b
            self._pending_init.append(len(idxs))
        self._pending = list(self._pending_init)
        self._sync = True
        self._works = []
        self._comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._events = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    # ------------------------------------------------------------------ hooks
    def _make_hook(self, i: int):
        def hook(param: torch.nn.Parameter):
            b = self._bucket_of[i]
            flat = self.buckets[b]
            g = param.grad
            if g is None or g.untyped_storage().data_ptr() != flat.untyped_storage().data_ptr():
                raise RuntimeError("a .grad was replaced (zero_grad(set_to_none=True)?): use GradAllReduce.zero_grad()")
            if not self._sync or self.world == 1:
                return
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b: int) -> None:
        flat = self.buckets[b]
        if self._comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))       # the bucket's gradients are complete here
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._works.append(work)
        else:
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # ------------------------------------------------------------------ API
    @contextmanager
    def no_sync(self):
        """Accumulate locally (non-final micro-batches; reference: `model.no_sync()`)."""
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    def finish(self) -> None:
        """Wait for the outstanding reductions and turn the sums into means.  Call after the last backward."""
        if self.world > 1:
            # buckets whose hooks did not all fire (unused parameters) are reduced now
            for b, left in enumerate(self._pending):
                if left != 0:
                    self._launch(b)
            for w in self._works:
                w.wait()            # NCCL: makes the current stream wait for the collective
            if self._comm_stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
            inv = 1.0 / self.world
            for flat in self.buckets:
                flat.mul_(inv)
        self._works = []
        self._pending = list(self._pending_init)

    def zero_grad(self) -> None:
        for flat in self.buckets:
            flat.zero_()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def bytes_per_reduce(self) -> int:
        return sum(b.numel() for b in self.buckets) * 4
