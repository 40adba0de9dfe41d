"""Model-level parity (GPU): the drop-in Swin-3D / ViT-3D modules against (a) the golden outputs of the
unmodified reference and (b) the CPU oracle's full gradient tensors, on identical synthetic weights/inputs.
Tolerance: 2e-2 relative (bf16 tensor-core path vs fp32 reference), as stated by BASELINE.json's north_star."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import swin3d_oracle as O  # noqa: E402
from oracle.cases import SWIN_CASES, VIT_CASES, SWIN_FULL, VIT_FULL, swin_ctor_kwargs, vit_ctor_kwargs  # noqa: E402
from oracle.synth import synth_volume, synth_targets, synth_keep_masks  # noqa: E402
from tests.helpers import meta, golden, rel_err, synth_sd  # noqa: E402

TOL = 2e-2


def _models():
    import vsn_b200  # noqa: F401
    from vsn_b200 import swin_model, vit_model
    return swin_model, vit_model


def _load_synth(model, shapes):
    sd = synth_sd(shapes, device="cuda")
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("relative_position_index" in k for k in missing)
    return sd


def _grad_report(model, ref_grads, g):
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        e = rel_err(p.grad, ref_grads[k])
        if e > worst[1]:
            worst = (k, e)
        # and against the reference's own numbers stored in the golden
        n = float(p.grad.double().norm())
        assert abs(n - float(g[f"gnorm/{k}"])) <= 2 * TOL * max(float(g[f"gnorm/{k}"]), 1e-6), (k, n, float(g[f"gnorm/{k}"]))
    return worst


@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_swin_matches_reference(name):
    swin_model, _ = _models()
    case, g, m = SWIN_CASES[name], golden(name), meta()[name]
    model = swin_model.SwinTransformerT(**swin_ctor_kwargs(case)).cuda()
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == m["state_shapes"]
    assert [k for k, _ in model.named_parameters()] == m["param_order"]
    _load_synth(model, m["state_shapes"])
    x = torch.from_numpy(synth_volume(case["input"], seed=1)).cuda()
    model.eval()
    with torch.no_grad():
        z = model(x)
        z16 = model(x.half())                    # fp16 volumes (the on-disk cache format) give the same result
    assert rel_err(z, g["logits_eval"]) < TOL
    assert rel_err(z16, g["logits_eval"]) < TOL
    # training step with injected DropPath decisions
    nblk = sum(case["depths"])
    masks = synth_keep_masks(max(2 * (nblk - 1), 1), x.shape[0], keep=0.7, seed=3)
    model.train()
    swin_model.DropPath.forced_masks = iter(torch.from_numpy(mm) for mm in masks)
    try:
        logits = model(x)
    finally:
        swin_model.DropPath.forced_masks = None
    tgt = torch.from_numpy(synth_targets(x.shape[0], case["num_classes"], seed=2)).cuda()
    loss = O.soft_target_ce(logits, tgt, 0.1)
    loss.backward()
    assert rel_err(logits.detach(), g["logits_train"]) < TOL
    assert abs(loss.item() - float(g["loss"])) < TOL * abs(float(g["loss"]))
    # oracle gradients (CPU fp32) on the same weights
    sd = synth_sd(m["state_shapes"])
    for v in sd.values():
        v.requires_grad_(True)
    zo = O.swin_forward(sd, x.cpu(), patch=case["patch_size"], window=case["window_size"], depths=case["depths"],
                        heads=case["num_heads"], drop_path_rate=case["drop_path"], training=True,
                        masks=iter(torch.from_numpy(mm) for mm in masks))
    O.soft_target_ce(zo, tgt.cpu(), 0.1).backward()
    worst = _grad_report(model, {k: sd[k].grad for k in m["param_order"]}, g)
    print("WORST_GRAD", name, worst)
    assert worst[1] < TOL, worst


@pytest.mark.parametrize("name", list(VIT_CASES))
def test_vit_matches_reference(name):
    _, vit_model = _models()
    case, g, m = VIT_CASES[name], golden(name), meta()[name]
    model = vit_model.ViTS(**vit_ctor_kwargs(case)).cuda()
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == m["state_shapes"]
    assert [k for k, _ in model.named_parameters()] == m["param_order"]
    model.load_state_dict(synth_sd(m["state_shapes"], device="cuda"))
    x = torch.from_numpy(synth_volume(case["input"], seed=1)).cuda()
    model.train()
    logits = model(x)
    assert rel_err(logits.detach(), g["logits_eval"]) < TOL
    tgt = torch.from_numpy(synth_targets(x.shape[0], case["num_classes"], seed=2)).cuda()
    loss = O.soft_target_ce(logits, tgt, 0.1)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < TOL * abs(float(g["loss"]))
    sd = synth_sd(m["state_shapes"])
    for v in sd.values():
        v.requires_grad_(True)
    zo = O.vit_forward(sd, x.cpu(), patch=case["patch_size"], heads=case["num_heads"], depth=case["depth"])
    O.soft_target_ce(zo, tgt.cpu(), 0.1).backward()
    worst = _grad_report(model, {k: sd[k].grad for k in m["param_order"]}, g)
    assert worst[1] < TOL, worst


def test_swin5c_full_size_seed0_logits():
    """Reference init under manual_seed(0) + the survey's input draws, full 144x168x144 volume."""
    swin_model, _ = _models()
    full = meta()["full"]
    torch.manual_seed(0)
    model = swin_model.SwinTransformerT(**swin_ctor_kwargs(dict(SWIN_FULL, num_classes=5, drop_path=0.15))).cuda().eval()
    gen = torch.Generator().manual_seed(1)
    x1 = torch.randn(1, 1, 144, 168, 144, generator=gen).half().float()
    x2 = torch.randn(1, 1, 50, 61, 47, generator=gen).half().float()
    with torch.no_grad():
        z1 = model(x1.cuda())[0]
        z2 = model(x2.cuda())[0]
    assert rel_err(z1, full["swin5c_seed0_in144x168x144"]) < TOL
    assert rel_err(z2, full["swin5c_seed0_in50x61x47"]) < TOL


def test_vit3c_full_size_seed0_logits():
    _, vit_model = _models()
    full = meta()["full"]
    torch.manual_seed(0)
    model = vit_model.ViTS(**vit_ctor_kwargs(dict(VIT_FULL, num_classes=3))).cuda().eval()
    gen = torch.Generator().manual_seed(1)
    torch.randn(1, 1, 144, 168, 144, generator=gen)
    torch.randn(1, 1, 50, 61, 47, generator=gen)
    x3 = torch.randn(1, 1, 144, 160, 144, generator=gen).half().float()
    with torch.no_grad():
        z = model(x3.cuda())[0]
    assert rel_err(z, full["vit3c_seed0_in144x160x144"]) < TOL


def test_in_kernel_gradient_accumulation_matches_autograd():
    """TrainStep sums the block gradients of an optimiser step's micro-batches inside the kernels
    (swin.GradAccumulation) and releases them to autograd once; same result as autograd's per-parameter adds."""
    swin_model, _ = _models()
    from vsn_b200 import swin
    from vsn_b200.train import TrainStep, soft_target_ce
    case = SWIN_CASES["swin_small_even"]
    model = swin_model.SwinTransformerT(**swin_ctor_kwargs(case)).cuda().train()
    _load_synth(model, meta()["swin_small_even"]["state_shapes"])
    xs = [torch.from_numpy(synth_volume(case["input"], seed=10 + i)).cuda() for i in range(3)]
    ys = [torch.from_numpy(synth_targets(xs[0].shape[0], case["num_classes"], seed=20 + i)).cuda() for i in range(3)]
    nblk = sum(case["depths"])

    def run(in_kernel):
        model.zero_grad(set_to_none=True)
        masks = synth_keep_masks(3 * max(2 * (nblk - 1), 1), xs[0].shape[0], keep=0.7, seed=5)
        swin_model.DropPath.forced_masks = iter(torch.from_numpy(mm) for mm in masks)
        try:
            if in_kernel:
                ts = TrainStep(model, use_ema=False)
                ts._accumulate(list(zip(xs, ys)))
            else:
                for x, y in zip(xs, ys):
                    (soft_target_ce(model(x), y, 0.1) / 3).backward()
        finally:
            swin_model.DropPath.forced_masks = None
        assert not swin.GradAccumulation.store and not swin.GradAccumulation.active
        return {k: p.grad.clone() for k, p in model.named_parameters()}

    ref, got = run(False), run(True)
    for k in ref:
        assert rel_err(got[k], ref[k]) < 1e-4, k


def test_droppath_factors_one_draw():
    """Training forward without forced masks: the DropPath factors of all blocks come from one Bernoulli draw and
    every factor is 0 or 1/keep of its block (timm DropPath semantics, models/swin_transformer_3d.py:251)."""
    swin_model, _ = _models()
    from vsn_b200 import swin
    case = SWIN_CASES["swin_small_even"]
    kw = dict(swin_ctor_kwargs(case))
    kw["stochastic_depth_prob"] = 0.3
    model = swin_model.SwinTransformerT(**kw).cuda().train()
    x = torch.from_numpy(synth_volume(case["input"], seed=3)).cuda()
    seen = []
    orig = swin.SwinBlockFn.apply

    def spy(*args):
        seen.append(args[-1])
        return orig(*args)

    swin.SwinBlockFn.apply = spy
    try:
        torch.manual_seed(0)
        model(x).sum().backward()
    finally:
        swin.SwinBlockFn.apply = orig
    probs = [blk.drop_path.drop_prob if isinstance(blk.drop_path, swin_model.DropPath) else 0.0
             for layer in model.backbone.layers for blk in layer.blocks]
    assert len(seen) == len(probs) and any(q > 0 for q in probs)
    for cfg, q in zip(seen, probs):
        if q == 0.0:
            assert cfg.scale1 is None and cfg.scale2 is None
            continue
        for sc in (cfg.scale1, cfg.scale2):
            assert sc.shape == (x.shape[0],) and sc.is_contiguous()
            ok = (sc == 0) | ((sc - 1.0 / (1.0 - q)).abs() < 1e-6)
            assert bool(ok.all())
