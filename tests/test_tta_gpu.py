"""Test-time augmentation + snapshot ensemble on the device (vsn_b200/tta.py, csrc/layout.cu:tta_views_f16_kernel)
against torch restatements of the reference's views (eval/test_time_augmentation.py:116-186,221-354): flip bit-exact,
crop + trilinear resize against `interpolate(align_corners=False)`, affine views against `grid_sample(bilinear, border)`,
and the batched prediction against the reference's view-by-view loop with its entropy weighting."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import cases  # noqa: E402


def _mods():
    import vsn_b200  # noqa: F401
    from vsn_b200 import ops, swin_model, tta
    return ops, swin_model, tta


def _volume(shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(2, 1, *shape, generator=g).half().cuda()


def test_views_flip_crop_affine_against_torch():
    ops, _, tta = _mods()
    shape = (20, 28, 24)
    x = _volume(shape)
    rot, tr = (0.04, -0.03, 0.05), (2.5, -4.0, 1.25)
    views = np.stack([tta.identity_view(shape), tta.flip_view(shape, 0), tta.center_crop_resize_view(shape, 0.9),
                      tta.affine_view(shape, rot, tr)])
    out = ops.tta_views(x, torch.from_numpy(views.astype(np.float32)).cuda()).view(2, 4, 1, *shape)
    assert torch.equal(out[:, 0], x)                                         # identity: bit-exact
    assert torch.equal(out[:, 1], torch.flip(x, dims=[2]))                   # RandFlip(prob=1, spatial_axis=0)
    # CenterSpatialCrop(int(size * 0.9)) + Resize(trilinear)  (:171-186,296-304)
    roi = [int(s * 0.9) for s in shape]
    st = [max(s // 2 - r // 2, 0) for s, r in zip(shape, roi)]
    crop = x[:, :, st[0]:st[0] + roi[0], st[1]:st[1] + roi[1], st[2]:st[2] + roi[2]].float()
    ref = F.interpolate(crop, size=shape, mode="trilinear", align_corners=False)
    assert float((out[:, 2].float() - ref).abs().max()) < 4e-3               # fp16 output of |x| < 5
    # RandAffine(bilinear, border): source = R (x - c) + c + t, sampled through grid_sample on the same coordinates
    m = torch.from_numpy(views[3, :12].reshape(3, 4)).float().cuda()
    dd, hh, ww = torch.meshgrid(*[torch.arange(s, device="cuda", dtype=torch.float32) for s in shape], indexing="ij")
    src = torch.stack([dd, hh, ww, torch.ones_like(dd)], -1) @ m.T           # [D,H,W,3] source voxel coordinates
    norm = torch.stack([2 * src[..., 2] / (shape[2] - 1) - 1, 2 * src[..., 1] / (shape[1] - 1) - 1,
                        2 * src[..., 0] / (shape[0] - 1) - 1], -1)           # grid_sample wants (x=w, y=h, z=d)
    ref = F.grid_sample(x.float(), norm[None].expand(2, -1, -1, -1, -1), mode="bilinear", padding_mode="border",
                        align_corners=True)
    assert float((out[:, 3].float() - ref).abs().max()) < 6e-3


def test_tta_prediction_equals_view_by_view_loop_and_ensemble_mean():
    ops, swin_model, tta = _mods()
    case = cases.SWIN_CASES["swin_small_even"]
    torch.manual_seed(0)
    model = swin_model.SwinTransformer(**cases.swin_ctor_kwargs(case)).cuda().eval()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, *case["input"][1:], generator=g)
    t = tta.TestTimeAugmentation(model, torch.device("cuda"), num_samples=2, seed=5)
    assert t.get_num_augmentations() == 5
    t2 = tta.TestTimeAugmentation(model, torch.device("cuda"), num_samples=2, seed=5)     # same draws
    got = t.predict(x)
    # the reference's loop (:246-354): one forward per view, entropy H = -sum p log p (p clamped at 1e-10),
    # weights 1 / (H + 1e-6) normalised over the views
    xv, V = t2.view_batch(x)
    xv = xv.view(2, V, *xv.shape[1:])
    for b in range(2):
        preds, ents = [], []
        for v in range(V):
            with torch.no_grad():
                p = torch.softmax(model(xv[b, v:v + 1]).float(), dim=1)
            preds.append(p)
            pc = torch.clamp(p.squeeze(), min=1e-10)
            ents.append(-torch.sum(pc * torch.log(pc)))
        wts = 1.0 / (torch.stack(ents) + 1e-6)
        wts = wts / wts.sum()
        ref = torch.sum(torch.stack(preds, 0).squeeze(1) * wts.unsqueeze(1), dim=0)
        assert float((got[b] - ref).abs().max()) < 2e-3
        assert abs(float(got[b].sum()) - 1.0) < 1e-5
    # snapshot ensemble: mean over the snapshots of the TTA prediction (scripts/transformer.sh:241-266)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    sd1 = {k: (v + 0.02 * torch.randn_like(v) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
    t3 = tta.TestTimeAugmentation(model, torch.device("cuda"), num_samples=2, seed=5)
    ens = tta.SnapshotEnsemble(model, [sd0, sd1], t3).predict(x)
    singles = []
    for sd in (sd0, sd1):
        model.load_state_dict(sd)
        singles.append(tta.TestTimeAugmentation(model, torch.device("cuda"), num_samples=2, seed=5).predict(x))
    assert float((ens - (singles[0] + singles[1]) / 2).abs().max()) < 1e-5
