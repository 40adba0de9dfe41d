"""Host-side pieces of the path that run without a GPU, against SURVEY.md's stage table and -- where the reference
checkout (or its installed copy under baseline/_ref) is present -- against the reference's own functions:
stage padding / window counts (models/swin_transformer_3d.py:457-461), the loss (regularization/label_smoothing.py:33-77),
the weight-decay groups (utils/helper.py:219-247) with the 65 / 108 split of Swin-T's 173 parameters, the multi-tensor
chunk table of the SAM / EMA / AdamW kernels, the shard map of the batch-sharded inference path."""
import os
import sys

import pytest
import torch

import vsn_b200  # noqa: F401
from vsn_b200 import ops, optim, swin, train

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_reference():
    from oracle import refshim
    try:
        return os.path.isdir(os.path.join(refshim.reference_root(), "models"))
    except Exception:      # noqa: BLE001
        return False


def test_stage_geometry_matches_the_survey_table():
    """swin-5c / swin-3c: input 144x168x144, patch 4, window (6,7,6): real grid -> padded grid, windows per volume."""
    window = (6, 7, 6)
    rows = [((36, 42, 36), (36, 42, 36), 216), ((18, 21, 18), (18, 21, 18), 27), ((9, 11, 9), (12, 14, 12), 8),
            ((5, 6, 5), (6, 7, 6), 1)]
    real = (36, 42, 36)
    for want_real, want_pad, want_windows in rows:
        assert real == want_real
        pad = swin.padded_dims(real, window)
        assert pad == want_pad
        g = ops.WindowGeom(2, pad, window, (3, 3, 3), True)
        assert g.N == 252 and g.S == 2 * want_windows
        assert list(g.arr) == [2, *pad, *window, 3, 3, 3, 1]
        real = tuple((r + 1) // 2 for r in real)          # PatchMerging pads odd dims by one and halves


def test_chunk_table_covers_every_element_once():
    sizes = [1, 65536, 65537, 0, 200000]
    chunk = 65536
    ct, co = optim.build_chunks(sizes, chunk)
    seen = {t: 0 for t in range(len(sizes))}
    for t, off in zip(ct, co):
        assert off == seen[t] and off < sizes[t]
        seen[t] = min(sizes[t], off + chunk)
    assert [seen[t] for t in range(len(sizes))] == sizes
    assert len(ct) == 1 + 1 + 2 + 0 + 4


def test_soft_target_ce_is_the_reference_loss():
    g = torch.Generator().manual_seed(0)
    z = torch.randn(6, 5, generator=g, dtype=torch.float64)
    y = torch.softmax(torch.randn(6, 5, generator=g, dtype=torch.float64) * 3, dim=1)      # MixUp-style soft labels
    for s in (0.0, 0.1):
        t = y * (1 - s) + s / 5
        want = -(t * torch.log_softmax(z, -1)).sum(-1).mean()
        assert torch.allclose(train.soft_target_ce(z, y, s), want, rtol=1e-12, atol=0)
    if _have_reference():
        from oracle import refshim
        refshim.install()
        try:
            from regularization.label_smoothing import LabelSmoothingLoss
            ref = LabelSmoothingLoss(smoothing=0.1)(z, y)
        finally:
            refshim.uninstall()
        assert torch.allclose(train.soft_target_ce(z, y, 0.1), ref, rtol=1e-12, atol=0)


def test_weight_decay_groups_follow_the_reference_rule():
    from vsn_b200.swin_model import SwinTransformerT
    m = SwinTransformerT(in_channels=1, patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                         window_size=[6, 7, 6], mlp_ratio=4.0, qkv_bias=True, dropout=0.0, attention_dropout=0.0,
                         stochastic_depth_prob=0.15, num_classes=5, norm_layer=torch.nn.LayerNorm)
    groups = train.param_groups(m)
    assert len(groups) == 2 and groups[1]["weight_decay"] == 0.0
    # Swin-T/5c has 173 parameter tensors, 65 weight-decayed (incl. the 12 bias tables), 108 not
    assert (len(groups[0]["params"]), len(groups[1]["params"])) == (65, 108)
    assert sum(p.numel() for grp in groups for p in grp["params"]) == 29_272_151
    names = {id(p): k for k, p in m.named_parameters()}
    assert all(not names[id(p)].endswith(".bias") and p.ndim > 1 for p in groups[0]["params"])
    assert sum("relative_position_bias_table" in names[id(p)] for p in groups[0]["params"]) == 12
    if _have_reference():
        from oracle import refshim
        refshim.install()
        try:
            from utils.helper import get_params_groups
            ref = get_params_groups(m)
        finally:
            refshim.uninstall()
        assert [[id(p) for p in grp["params"]] for grp in ref] == [[id(p) for p in grp["params"]] for grp in groups]


def test_relative_position_index_closed_form():
    """idx(i, j) = lin(i) - lin(j) + offset with lin over the (2W-1) box (models/swin_transformer_3d.py:132-152)."""
    wd, wh, ww = 6, 7, 6
    rpi = swin.relative_position_index((wd, wh, ww))
    assert rpi.shape == (252, 252) and rpi.dtype == torch.int64
    assert int(rpi.min()) == 0 and int(rpi.max()) == (2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1) - 1 == 1572
    t = torch.arange(252)
    d, h, w = t // (wh * ww), (t // ww) % wh, t % ww
    lin = d * ((2 * wh - 1) * (2 * ww - 1)) + h * (2 * ww - 1) + w
    off = (wd - 1) * (2 * wh - 1) * (2 * ww - 1) + (wh - 1) * (2 * ww - 1) + (ww - 1)
    assert torch.equal(rpi, lin[:, None] - lin[None, :] + off)
