"""Host-side logic of the batch-sharded inference path (vsn_b200/tta.py: ShardedInference, BASELINE config 5 on
several GPUs) on CPU: world size 2 and 3 over gloo with a stand-in predictor.  Every rank must end up
with the full [N, K] table in subject order, equal to the single-process loop, for N above, equal to and below the
world size (ranks without subjects still join the one all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vsn_b200  # noqa: F401
from vsn_b200.tta import ShardedInference, shard_range


def _predictor(x):
    """Deterministic stand-in for TestTimeAugmentation / SnapshotEnsemble: [b,1,D,H,W] -> [b,4] probabilities."""
    x = x.float()
    f = torch.stack([x.mean((1, 2, 3, 4)), x.amax((1, 2, 3, 4)), x.amin((1, 2, 3, 4)), x.std((1, 2, 3, 4))], dim=1)
    return torch.softmax(f, dim=1)


def _volumes(n):
    g = torch.Generator().manual_seed(7)
    return torch.randn(n, 1, 4, 5, 3, generator=g).half()


def test_shard_range_is_a_balanced_contiguous_partition():
    for n in (0, 1, 2, 5, 16, 17):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_single_process_is_the_plain_loop():
    v = _volumes(5)
    assert torch.equal(ShardedInference(_predictor, batch=2).predict(v), _predictor(v))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for n in (7, world, 1):                       # ragged shards, one subject each, fewer subjects than ranks
            v = _volumes(n)
            sh = ShardedInference(_predictor, batch=2)
            assert (sh.world, sh.rank) == (world, rank)
            local, (lo, hi) = sh.predict_local(v)
            assert local.shape[0] == hi - lo
            full = sh.gather(local, n)
            want = _predictor(v)
            ok = ok and full.shape == want.shape and bool(torch.allclose(full, want, rtol=0, atol=1e-7))
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_inference_gloo(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: 1 for r in range(world)}
