"""The Python-level drop-in boundary of SURVEY.md 8(b): what the reference's callers hand to `forward` and to the
constructors, beyond the plain contiguous fp32 batch the parity tests use.

CPU part: every option combination the kernels do not implement raises loudly at construction ("never silently differ";
models/swin_transformer_3d.py:702-726, models/vit_3d.py:460-507), a CPU tensor raises (no CPU fallback).
GPU part: the forms the trainer / evaluation script actually produce -- channels_last_3d strides
(train/train_transformer.py:1126, eval/eval_transformer.py:441), fp16 volumes, a Tensor subclass (MONAI MetaTensor stand-in),
torch.inference_mode, torch.autocast, a module converted with .to(memory_format=channels_last_3d) (:2093), batch 1 --
all give the result of the plain call bit for bit (the kernels own layout and operand type)."""
import pytest
import torch

from oracle.cases import SWIN_CASES, VIT_CASES, swin_ctor_kwargs, vit_ctor_kwargs
from oracle.synth import synth_volume
from tests.helpers import meta, synth_sd


def _models():
    import vsn_b200  # noqa: F401
    from vsn_b200 import swin_model, vit_model
    return swin_model, vit_model


# ------------------------------------------------------------------------------------------------- CPU: loud raises
@pytest.mark.parametrize("override", [
    dict(post_norm=True), dict(layer_scale=True), dict(enable_stable=True), dict(dropout=0.1), dict(attention_dropout=0.1),
    dict(norm_layer=torch.nn.BatchNorm1d), dict(use_shakedrop=True, stochastic_depth_prob=0.2), dict(in_channels=2),
    dict(embed_dim=40, num_heads=[1, 2]),        # head_dim 40: no kernel
])
def test_swin_unsupported_options_raise_at_construction(override):
    swin_model, _ = _models()
    kw = swin_ctor_kwargs(SWIN_CASES["swin_small_even"])
    kw.update(override)
    with pytest.raises((NotImplementedError, ValueError, RuntimeError)):
        swin_model.SwinTransformerT(**kw)


@pytest.mark.parametrize("override", [dict(post_norm=True), dict(dropout=0.1), dict(attention_dropout=0.1),
                                      dict(layer_scale=True), dict(dim_head=48)])
def test_vit_unsupported_options_raise_at_construction(override):
    _, vit_model = _models()
    kw = vit_ctor_kwargs(VIT_CASES["vit_tiny"])
    kw.update(override)
    with pytest.raises((NotImplementedError, ValueError, RuntimeError)):
        vit_model.ViTS(**kw)


def test_supported_defaults_construct_and_cpu_input_raises():
    swin_model, vit_model = _models()
    m = swin_model.SwinTransformerT(**swin_ctor_kwargs(SWIN_CASES["swin_small_even"]), use_checkpoint=True)
    assert getattr(m, "use_checkpoint", None) is not None or True      # accepted (read via getattr by the trainer)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 48, 56, 48))
    v = vit_model.ViTS(**vit_ctor_kwargs(VIT_CASES["vit_tiny"]))
    with pytest.raises(RuntimeError, match="CUDA"):
        v(torch.zeros(1, 1, 32, 48, 32))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 48, 56, 48))                                   # not [B,1,D,H,W]


# ------------------------------------------------------------------------------------------------- GPU: input forms
class _Meta(torch.Tensor):
    """Stand-in for monai.data.MetaTensor: a Tensor subclass that carries extra attributes."""
    @staticmethod
    def __new__(cls, x, *a, **kw):
        return torch.Tensor._make_subclass(cls, x)


def _forms(x):
    yield "channels_last_3d", x.to(memory_format=torch.channels_last_3d)
    yield "fp16", x.half()
    yield "subclass", _Meta(x.clone())
    big = torch.zeros(x.shape[0], 1, x.shape[2], x.shape[3], x.shape[4] * 2, device=x.device)
    big[..., ::2] = x
    yield "strided_view", big[..., ::2]


def _check_forms(model, x):
    model.eval()
    with torch.no_grad():
        ref = model(x)
        assert type(ref) is torch.Tensor and ref.dtype == torch.float32
        for name, xf in _forms(x):
            z = model(xf)
            assert type(z) is torch.Tensor, name
            assert torch.equal(z, ref), (name, float((z - ref).abs().max()))
        with torch.autocast("cuda", dtype=torch.float16):
            za = model(x)
        assert za.dtype == torch.float32 and torch.equal(za, ref), "autocast (the kernels own the operand type)"
        z1 = model(x[:1])                                 # another batch size: another tile schedule, same sample result
        assert float((z1 - ref[:1]).norm() / ref[:1].norm()) < 1e-5, "batch 1"
    with torch.inference_mode():
        zi = model(x)
    assert torch.equal(zi, ref), "inference_mode"
    # the trainer converts the module itself (5-D conv weight takes channels_last_3d strides) and DDP-style wrappers
    # read parameters as leaves: same result, every parameter still a leaf that receives a gradient
    model = model.to(memory_format=torch.channels_last_3d)
    with torch.no_grad():
        assert torch.equal(model(x), ref), "module in channels_last_3d"
    model.train()
    model(x).sum().backward()
    for k, p in model.named_parameters():
        assert p.is_leaf and p.grad is not None and p.grad.shape == p.shape, k


@pytest.mark.gpu
def test_swin_forward_accepts_the_callers_input_forms():
    swin_model, _ = _models()
    name = "swin_tiny_odd"
    case, m = SWIN_CASES[name], meta()[name]
    model = swin_model.SwinTransformerT(**swin_ctor_kwargs(case, drop_path=0.0)).cuda()
    model.load_state_dict(synth_sd(m["state_shapes"], device="cuda"), strict=False)
    x = torch.from_numpy(synth_volume(case["input"], seed=1)).half().float().cuda()     # exactly representable in fp16
    _check_forms(model, x)


@pytest.mark.gpu
def test_vit_forward_accepts_the_callers_input_forms():
    _, vit_model = _models()
    name = "vit_tiny"
    case, m = VIT_CASES[name], meta()[name]
    model = vit_model.ViTS(**vit_ctor_kwargs(case)).cuda()
    model.load_state_dict(synth_sd(m["state_shapes"], device="cuda"), strict=False)
    x = torch.from_numpy(synth_volume(case["input"], seed=1)).half().float().cuda()
    _check_forms(model, x)
