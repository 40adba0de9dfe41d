"""Host-side logic of the data-parallel gradient exchange (vsn_b200/ddp.py) on CPU: world size 2 over gloo.
Covers bucket planning, the no_sync accumulation schedule of the reference's micro-batch loop
(train/train_transformer.py:1131-1137), the mean over ranks and gradient views surviving optimiser steps."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vsn_b200  # noqa: F401
from vsn_b200.ddp import GradAllReduce, plan_buckets


def test_plan_buckets_reverse_order_and_caps():
    sizes = [10, 20, 30, 5, 100, 1]
    plan = plan_buckets(sizes, cap_elems=40)
    assert [i for b in plan for i in b] == [5, 4, 3, 2, 1, 0]          # reverse registration order
    assert plan == [[5], [4], [3, 2], [1, 0]]                          # oversize parameter alone, cap respected
    assert all(sum(sizes[i] for i in b) <= 40 or len(b) == 1 for b in plan)
    assert plan_buckets([7], 1) == [[0]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8), torch.nn.LayerNorm(8),
                               torch.nn.Linear(8, 3))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, ref = _model(), _model()
        ddp = GradAllReduce(model.parameters(), bucket_mb=0.001)       # ~262 elements per bucket: several buckets
        assert len(ddp.buckets) > 1
        g = torch.Generator().manual_seed(100 + rank)
        micro = [(torch.randn(4, 16, generator=g), torch.randn(4, 3, generator=g)) for _ in range(3)]
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        for step in range(2):
            # reference schedule: all but the last micro-batch under no_sync
            for i, (x, y) in enumerate(micro):
                if i < len(micro) - 1:
                    with ddp.no_sync():
                        ((model(x) - y) ** 2).mean().backward()
                else:
                    ((model(x) - y) ** 2).mean().backward()
            ddp.finish()
            # expectation: mean over ranks of the locally accumulated gradients
            ref.load_state_dict(model.state_dict())
            ref.zero_grad(set_to_none=True)
            for x, y in micro:
                ((ref(x) - y) ** 2).mean().backward()
            for p, q in zip(model.parameters(), ref.parameters()):
                want = q.grad.clone()
                dist.all_reduce(want)
                want /= world
                assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-6), (rank, step)
            ptrs = [p.grad.data_ptr() for p in model.parameters()]
            opt.step()
            ddp.zero_grad()
            assert ptrs == [p.grad.data_ptr() for p in model.parameters()]   # views survive the step
            assert all(float(p.grad.abs().sum()) == 0.0 for p in model.parameters())
        # weights stay identical across ranks
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_grad_allreduce_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
