"""Pin the CPU oracle (oracle/swin3d_oracle.py) against goldens generated from the
unmodified reference (oracle/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import swin3d_oracle as O
from oracle.cases import SWIN_CASES, VIT_CASES
from oracle.synth import synth_volume, synth_targets, synth_keep_masks
from tests.helpers import meta, golden, rel_err, synth_sd


def sha16(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_relative_position_index_kat():
    k = meta()["kat"]["rpi"]
    rpi = O.relative_position_index((6, 7, 6))
    assert rpi.dtype == np.int64 and list(rpi.shape) == k["shape"]
    assert (int(rpi.min()), int(rpi.max()), int(rpi.sum())) == (k["min"], k["max"], k["sum"])
    assert rpi[0, :8].tolist() == k["row0"] and int(rpi[251, 0]) == k["last0"]
    assert sha16(rpi) == k["sha16"]
    assert sha16(O.relative_position_index((7, 7, 7))) == meta()["kat"]["rpi_777"]["sha16"]


@pytest.mark.parametrize("tag", ["stage0", "stage1", "stage2", "stage3", "odd0", "odd1"])
def test_shift_mask_bit_exact(tag):
    k = meta()["kat"]["mask"][tag]
    grid = O.padded_grid(k["real"], (6, 7, 6))
    m = O.shift_mask(grid, (6, 7, 6), (3, 3, 3))
    assert list(m.shape) == k["shape"] and m.dtype == np.float32
    assert int((m != 0).sum()) == k["nonzero"]
    assert sorted(set(np.unique(m).tolist())) == k["values"]
    assert sha16(m) == k["sha16"]


def test_window_partition_order():
    k = meta()["kat"]["window_partition_12x14x12"]
    t = O.window_tokens((12, 14, 12), (6, 7, 6))
    assert sha16(t) == k["sha16"] and t[3, :8].tolist() == k["w3_head"]
    # window_reverse(window_partition(x)) == x : the table is a permutation
    assert sorted(t.reshape(-1).tolist()) == list(range(12 * 14 * 12))


def _check_grads(g, grads, tol):
    worst = 0.0
    for k, v in grads.items():
        n = float(v.double().norm())
        assert abs(n - float(g[f"gnorm/{k}"])) <= tol * max(float(g[f"gnorm/{k}"]), 1e-6), k
        if f"gfull/{k}" in g:
            worst = max(worst, rel_err(v, g[f"gfull/{k}"]))
        np.testing.assert_allclose(v.reshape(-1)[:8].numpy(), g[f"ghead/{k}"], rtol=50 * tol, atol=1e-5)
    assert worst <= tol, worst


@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_swin_oracle_matches_reference(name):
    case, g, m = SWIN_CASES[name], golden(name), meta()[name]
    sd = synth_sd(m["state_shapes"])
    x = torch.from_numpy(synth_volume(case["input"], seed=1))
    kw = dict(patch=case["patch_size"], window=case["window_size"], depths=case["depths"], heads=case["num_heads"])
    taps = {}
    with torch.no_grad():
        z = O.swin_forward(sd, x, **kw, taps=taps)
    assert rel_err(z, g["logits_eval"]) < 2e-5
    for k, v in taps.items():
        assert abs(float(v.double().norm()) - float(g[f"tapnorm/{k}"])) < 2e-5 * float(g[f"tapnorm/{k}"]), k
    for v in sd.values():
        v.requires_grad_(True)
    nblk = sum(case["depths"])
    masks = synth_keep_masks(max(2 * (nblk - 1), 1), x.shape[0], keep=0.7, seed=3)
    z = O.swin_forward(sd, x, **kw, drop_path_rate=case["drop_path"], training=True,
                       masks=iter(torch.from_numpy(mm) for mm in masks))
    tgt = torch.from_numpy(synth_targets(x.shape[0], case["num_classes"], seed=2))
    loss = O.soft_target_ce(z, tgt, 0.1)
    loss.backward()
    assert rel_err(z.detach(), g["logits_train"]) < 2e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    _check_grads(g, {k: sd[k].grad for k in m["param_order"]}, 2e-4)


@pytest.mark.parametrize("name", list(VIT_CASES))
def test_vit_oracle_matches_reference(name):
    case, g, m = VIT_CASES[name], golden(name), meta()[name]
    sd = synth_sd(m["state_shapes"])
    for v in sd.values():
        v.requires_grad_(True)
    x = torch.from_numpy(synth_volume(case["input"], seed=1))
    z = O.vit_forward(sd, x, patch=case["patch_size"], heads=case["num_heads"], depth=case["depth"])
    assert rel_err(z.detach(), g["logits_eval"]) < 2e-5
    tgt = torch.from_numpy(synth_targets(x.shape[0], case["num_classes"], seed=2))
    loss = O.soft_target_ce(z, tgt, 0.1)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    _check_grads(g, {k: sd[k].grad for k in m["param_order"]}, 2e-4)


def test_sam_oracle_matches_reference():
    g = golden("sam")
    p0 = [g[f"p0_{i}"] for i in range(4)]
    g0 = [g[f"g0_{i}"] for i in range(4)]
    for adaptive in (False, True):
        a = int(adaptive)
        n = O.sam_grad_norm(g0, p0, adaptive)
        assert abs(n - float(g[f"norm_{a}"])) < 1e-6 * n
        pert, old = O.sam_first_step([p.copy() for p in p0], g0, 0.05, adaptive)
        for i in range(4):
            np.testing.assert_allclose(pert[i], g[f"pert_{a}_{i}"], rtol=1e-6, atol=1e-7)
            np.testing.assert_array_equal(old[i], p0[i])


def test_ema_oracle_matches_reference():
    g = golden("ema")
    keys = sorted({k.split("/", 1)[1] for k in g if k.startswith("s0/")})
    for step in range(1, 6):
        lo = max(0, step - 2)
        for k in keys:
            states = [g[f"s{j}/{k}"] for j in range(lo, step + 1)]
            want = g[f"ema{step}/{k}"]
            if np.issubdtype(want.dtype, np.floating):
                np.testing.assert_allclose(O.ema_average(states, 0.999), want, rtol=1e-6, atol=1e-7)
            else:
                np.testing.assert_array_equal(states[-1], want)


def test_product_index_buffer_and_seed0_init_bit_exact():
    """The drop-in's `relative_position_index` buffer (state_dict payload) and its seed-0 initialisation are
    bit-identical to the reference's (models/swin_transformer_3d.py:132-152,159,676-683): sha1 of the bytes against
    the hashes the golden generator took from the unmodified reference.  Host-side only (no kernel runs)."""
    import hashlib
    import vsn_b200  # noqa: F401
    from vsn_b200 import swin, swin_model, vit_model
    from oracle.cases import SWIN_FULL, VIT_FULL, swin_ctor_kwargs, vit_ctor_kwargs

    def sha16(a):
        return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]

    kat, full = meta()["kat"], meta()["full"]
    rpi = swin.relative_position_index((6, 7, 6)).numpy()
    assert rpi.dtype == np.int64 and list(rpi.shape) == kat["rpi"]["shape"] and sha16(rpi) == kat["rpi"]["sha16"]
    assert sha16(swin.relative_position_index((7, 7, 7)).numpy()) == kat["rpi_777"]["sha16"]
    torch.manual_seed(0)
    m5 = swin_model.SwinTransformerT(**swin_ctor_kwargs(dict(SWIN_FULL, num_classes=5, drop_path=0.15)))
    sd = m5.state_dict()
    assert len(sd) == full["swin5c_n_keys"] and sum(p.numel() for p in m5.parameters()) == full["swin5c_n_params"]
    assert {k: list(v.shape) for k, v in sd.items()} == full["swin5c_state_shapes"]
    for k, h in full["swin5c_param_sha16"].items():
        assert sha16(sd[k].numpy()) == h, k
    assert sha16(sd["backbone.layers.1.blocks.0.attn.relative_position_index"].numpy()) == kat["rpi"]["sha16"]
    torch.manual_seed(0)
    mv = vit_model.ViTS(**vit_ctor_kwargs(dict(VIT_FULL, num_classes=3)))
    sdv = mv.state_dict()
    assert {k: list(v.shape) for k, v in sdv.items()} == full["vit3c_state_shapes"]
    for k, h in full["vit3c_param_sha16"].items():
        assert sha16(sdv[k].numpy()) == h, k
