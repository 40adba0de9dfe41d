"""The input step on the device (vsn_b200/data.py + csrc/layout.cu: MixUp + whole-image statistics + z-score on fp16
volumes) against tests/golden/mixup.npz -- outputs of the UNMODIFIED reference's MRIMixUp followed by
NormalizeIntensity() (dataset/dataset.py:186-286, train/train_transformer.py:1729-1752)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.helpers import golden  # noqa: E402

NAMES = ["AD", "CN", "FTD"]


def _mods():
    import vsn_b200  # noqa: F401
    from vsn_b200 import data, ops
    return data, ops


@pytest.mark.parametrize("epoch", [0, 3])
def test_device_pipeline_matches_reference_mixup_and_normalize(epoch):
    data, _ = _mods()
    g = golden("mixup")
    vols = torch.from_numpy(g["volumes"]).pin_memory()
    labels = torch.from_numpy(g["labels"])
    pipe = data.DeviceInputPipeline(vols, labels, [NAMES[int(i)] for i in g["diagnoses"]], alpha=0.3, mixup_prob=0.7, seed=7)
    idx = list(range(vols.shape[0]))
    x, y = pipe.batch(idx, epoch=epoch)
    assert x.dtype == torch.float16 and x.shape == vols.shape and x.is_cuda
    ref = g[f"x_e{epoch}"]
    # fp16 output of values ~N(0,1) (|x| < 8: half an ulp is 2e-3) + one fp16 ulp of the RAW mixed voxel where the
    # two-step rounding order of the in-place fp16 MixUp differs between CPU and GPU arithmetic, scaled by 1/std
    np.testing.assert_allclose(x.float().cpu().numpy(), ref, rtol=0, atol=6e-3)
    assert float(np.abs(x.float().cpu().numpy() - ref).mean()) < 6e-4
    np.testing.assert_allclose(y.cpu().numpy(), g[f"y_e{epoch}"], rtol=1e-6, atol=1e-7)
    # the constant volume (std == 0) is only centred; unless it was mixed it comes out exactly zero
    if int(g[f"partner_e{epoch}"][5]) < 0:
        assert float(x[5].abs().max()) == 0.0


def test_volume_stats_and_zscore_full_size():
    """[B,1,144,168,144] fp16 (BASELINE.json's volume): mean / population std against float64 numpy; the normalised batch
    has mean 0 and std 1; a batch of a different size re-uses nothing stale (the scratch sums are cleared)."""
    _, ops = _mods()
    g = torch.Generator().manual_seed(3)
    for B in (3, 2):
        x = (torch.rand(B, 1, 144, 168, 144, generator=g) * 700.0).half()
        x[0, :, :72] = 0                                         # skull-stripped background
        xd = x.cuda()
        stats = ops.volume_stats(xd).cpu().numpy()
        for b in range(B):
            v = x[b].double().numpy()
            assert abs(stats[b, 0] - v.mean()) <= 1e-5 * abs(v.mean())
            assert abs(1.0 / stats[b, 1] - v.std()) <= 1e-5 * v.std()
        out, _ = ops.mixup_zscore(xd)
        o = out.float()
        assert float(o.mean(dim=(1, 2, 3, 4)).abs().max()) < 2e-3
        assert float((o.std(dim=(1, 2, 3, 4), unbiased=False) - 1).abs().max()) < 2e-3


def test_mixup_matches_two_step_fp16_rounding():
    """dataset/dataset.py:276-277: `sample1.mul_(alpha)` rounds to fp16, `.add_(sample2, alpha=1 - alpha)` rounds again."""
    _, ops = _mods()
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(4, 1, 16, 16, 16, generator=g) * 900.0).half()
    lam = torch.tensor([0.3, 1.0, 0.71, 0.05])
    perm = torch.tensor([2, 1, 0, 1], dtype=torch.int32)
    out = ops.mixup(x.cuda(), lam.cuda(), perm.cuda()).cpu()
    ref = x.clone()
    for b in range(4):
        if float(lam[b]) != 1.0:
            ref[b] = x[b].clone().mul_(float(lam[b])).add_(x[int(perm[b])], alpha=1.0 - float(lam[b]))
    d = (out.float() - ref.float()).abs()
    ulp = torch.maximum(ref.float().abs(), torch.tensor(1.0)) * 2.0 ** -10
    assert bool((d <= ulp).all()) and float((d == 0).float().mean()) > 0.999     # bit-exact but for fma contraction
