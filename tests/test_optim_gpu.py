"""SAM and EMA on the CUDA multi-tensor kernels (vsn_b200/optim.py + csrc/optim.cu) against the goldens the
UNMODIFIED reference produced (tests/golden/sam.npz, ema.npz: regularization/sam.py:38-155, utils/ema.py:72-142).
fp32 element-wise arithmetic: tolerance 1e-6 relative (fma contraction / summation order only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.helpers import golden  # noqa: E402

SHAPES = 4


def _optim():
    import vsn_b200  # noqa: F401
    from vsn_b200 import optim
    return optim


def _params(g):
    return [torch.nn.Parameter(torch.from_numpy(g[f"p0_{i}"].copy()).cuda()) for i in range(SHAPES)]


def _close(a, b, rtol=1e-6, atol=1e-7):
    np.testing.assert_allclose(a.detach().cpu().numpy(), b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("adaptive", [False, True])
def test_sam_two_pass_step_matches_reference(adaptive):
    optim = _optim()
    g = golden("sam")
    a = int(adaptive)
    ps = _params(g)
    opt = optim.SAM([{"params": ps[:2]}, {"params": ps[2:], "weight_decay": 0.0}], torch.optim.AdamW, rho=0.05,
                    adaptive=adaptive, lr=1e-3, weight_decay=0.05)
    assert isinstance(opt.base_optimizer, torch.optim.AdamW) and opt.param_groups is opt.base_optimizer.param_groups
    for i, p in enumerate(ps):
        p.grad = torch.from_numpy(g[f"g0_{i}"].copy()).cuda()
    n = float(opt._grad_norm())
    assert abs(n - float(g[f"norm_{a}"])) < 1e-6 * n
    opt.first_step(zero_grad=True)                                   # w -> w + e(w), old_p saved, grads dropped
    assert all(p.grad is None for p in ps)
    for i, p in enumerate(ps):
        _close(p, g[f"pert_{a}_{i}"])
        _close(opt.state[p]["old_p"], g[f"p0_{i}"], rtol=0, atol=0)   # the saved weights are bit-exact copies
    for i, p in enumerate(ps):
        p.grad = torch.from_numpy((0.5 * g[f"g0_{i}"]).copy()).cuda()
    opt.second_step(zero_grad=True)                                   # back to w, AdamW step with the new grads
    for i, p in enumerate(ps):
        _close(p, g[f"final_{a}_{i}"], rtol=2e-6)


@pytest.mark.parametrize("tag", ["inf1", "nan2", "allbad", "zero"])
def test_sam_non_finite_and_zero_gradients_follow_reference(tag):
    """Tensors with a non-finite norm stay out of the global norm and are not perturbed; an all-zero gradient disables
    the perturbation (regularization/sam.py:46-52,66-70,141-153).  In every case second_step restores the weights
    exactly (the reference returns early without saving old_p when the norm is zero; here old_p is always saved)."""
    optim = _optim()
    g = golden("sam")
    ps = _params(g)
    opt = optim.SAM([{"params": ps}], torch.optim.SGD, rho=0.05, adaptive=False, lr=0.0)
    grads = [g[f"g0_{i}"].copy() for i in range(SHAPES)]
    if tag == "inf1":
        grads[1][3] = np.inf
    elif tag == "nan2":
        grads[2][0, 1, 2] = np.nan
    elif tag == "allbad":
        for x in grads:
            x.reshape(-1)[0] = np.inf
    else:
        grads = [np.zeros_like(x) for x in grads]
    for p, x in zip(ps, grads):
        p.grad = torch.from_numpy(x).cuda()
    n = float(opt._grad_norm())
    want = float(g[f"nf_{tag}_norm"])
    assert abs(n - want) <= 1e-6 * max(want, 1e-12), (n, want)
    opt.first_step(zero_grad=False)
    for i, p in enumerate(ps):
        _close(p, g[f"nf_{tag}_pert_{i}"])
        assert torch.isfinite(p).all()
    for p, x in zip(ps, grads):
        p.grad = torch.zeros_like(p)
    opt.second_step(zero_grad=True)                                   # lr = 0: only the restore is visible
    for i, p in enumerate(ps):
        _close(p, g[f"p0_{i}"], rtol=0, atol=0)


def test_sam_plan_survives_load_state_dict():
    """optimizer.load_state_dict replaces optimizer.state (new old_p tensors): the cached device pointer tables must
    follow, or first_step would save the weights into freed memory (round-1 advisor finding)."""
    optim = _optim()
    g = golden("sam")
    ps = _params(g)
    opt = optim.SAM([{"params": ps}], torch.optim.SGD, rho=0.05, adaptive=False, lr=0.0)

    def one_pass():
        for i, p in enumerate(ps):
            p.grad = torch.from_numpy(g[f"g0_{i}"].copy()).cuda()
        opt.first_step(zero_grad=False)
        for i, p in enumerate(ps):
            _close(p, g[f"pert_0_{i}"])
        opt.second_step(zero_grad=True)
        for i, p in enumerate(ps):
            _close(p, g[f"p0_{i}"], rtol=0, atol=0)

    one_pass()
    import io
    buf = io.BytesIO()
    torch.save(opt.state_dict(), buf)                                 # a checkpoint round trip: fresh state tensors
    buf.seek(0)
    olds = [opt.state[p]["old_p"] for p in ps]
    opt.load_state_dict(torch.load(buf, map_location="cpu"))
    assert all(opt.state[p]["old_p"] is not o for p, o in zip(ps, olds))
    del olds
    junk = [torch.full((1 << 16,), float("nan"), device="cuda") for _ in range(8)]   # reuse freed allocator blocks
    one_pass()
    for p in ps:
        assert opt.state[p]["old_p"].data_ptr() in {int(x) for x in opt._plan["old"].tolist()}
    del junk


def test_ema_ring_matches_reference_and_apply_restore_round_trip():
    """Five updates: the 1 -> 2 -> 3 snapshot ramp, then two more that recycle ring slots; weights
    decay^(k-1-i)/sum, integer buffers follow the newest snapshot (utils/ema.py:72-108).  Then apply_to / restore
    (utils/ema.py:110-142) swap the average in and the live weights back, bit for bit."""
    optim = _optim()
    g = golden("ema")
    keys = sorted({k.split("/", 1)[1] for k in g if k.startswith("s0/")})
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3)).cuda()   # as the golden generator's
    live = net.state_dict()
    assert sorted(live) == keys
    with torch.no_grad():
        for k in keys:
            live[k].copy_(torch.from_numpy(g[f"s0/{k}"]).cuda())
    ema = optim.EMAModel(model=net, decay=0.999)
    for k in keys:
        _close(ema.model_state[k], g[f"s0/{k}"], rtol=0, atol=0)     # a single snapshot: the average is the state
    for step in range(1, 6):
        with torch.no_grad():
            for k in keys:
                live[k].copy_(torch.from_numpy(g[f"s{step}/{k}"]).cuda())
        ema.update(net)
        for k in keys:
            want = g[f"ema{step}/{k}"]
            if np.issubdtype(want.dtype, np.floating):
                _close(ema.model_state[k], want)
            else:
                assert ema.model_state[k].cpu().numpy().tolist() == want.tolist()
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ema.apply_to(net)
    for k in keys:
        assert torch.equal(net.state_dict()[k], ema.model_state[k]), k
    ema.restore(net)
    for k in keys:
        assert torch.equal(net.state_dict()[k], before[k]), k
    assert ema.orig_state is None
