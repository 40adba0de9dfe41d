"""Parity numbers of one precision mode, printed as one JSON line (run as a script: the precision of the 16-bit operands
is fixed per process by VSN_B200_PRECISION, so tests/test_precision_f16_gpu.py runs this file in a child process).

For every small Swin / ViT case and for the two full-size train goldens: relative error of the eval logits, the
train-mode logits and the loss against the goldens of the unmodified reference, and the worst per-parameter gradient
error against the CPU oracle (small cases, whole tensors) or the golden samples (full size).  The backward runs under a
loss scale (as the reference's GradScaler does, train/train_transformer.py:1141-1160): half-precision gradients
underflow without one; the reported gradients are unscaled in fp32.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import swin3d_oracle as O  # noqa: E402
from oracle.cases import SWIN_CASES, VIT_CASES, SWIN_FULL, VIT_FULL, swin_ctor_kwargs, vit_ctor_kwargs  # noqa: E402
from oracle.synth import synth_volume, synth_targets, synth_keep_masks  # noqa: E402
from tests.helpers import meta, golden, rel_err, synth_sd  # noqa: E402


def _worst_small(model, ref_grads, scale):
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        e = rel_err(p.grad.float() / scale, ref_grads[k])
        if e > worst[1]:
            worst = (k, e)
    return worst


def _worst_full(model, g, scale):
    from oracle.make_golden import sample_index
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        gr = p.grad.detach().float().reshape(-1) / scale
        idx = torch.from_numpy(sample_index(k, gr.numel())).cuda()
        e = rel_err(gr[idx], g[f"gsamp/{k}"])
        if f"gfull/{k}" in g:
            e = max(e, rel_err(gr, g[f"gfull/{k}"]))
        if e > worst[1]:
            worst = (k, e)
    return worst


def reference_modes(scale):
    """The UNMODIFIED reference on the same GPU in its own reduced-precision modes -- TF32 matmuls
    (utils/seed.py:50-51, train/train_transformer.py:91-92) and float16 autocast (:1141-1160) -- against the same fp32
    goldens / oracle gradients: how far the reference's own paths sit from its fp32 numbers (context for the f16
    mode's figures; test infrastructure, nothing of this repo's product runs here)."""
    from oracle import refshim
    refshim.install()
    from models.swin_transformer_3d import SwinTransformerT
    from models.vit_3d import ViTS
    from timm.layers import DropPath
    res = {}
    for mode in ("tf32", "fp16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = mode == "tf32"
        res[mode] = {}
        for name, case in list(SWIN_CASES.items()) + list(VIT_CASES.items()):
            is_swin = name in SWIN_CASES
            g, m = golden(name), meta()[name]
            model = (SwinTransformerT(**swin_ctor_kwargs(case)) if is_swin else ViTS(**vit_ctor_kwargs(case))).cuda()
            model.load_state_dict(synth_sd(m["state_shapes"], device="cuda"), strict=False)
            x = torch.from_numpy(synth_volume(case["input"], seed=1)).cuda()
            masks = None
            if is_swin:
                masks = synth_keep_masks(max(2 * (sum(case["depths"]) - 1), 1), x.shape[0], keep=0.7, seed=3)
                DropPath.forced_masks = iter(torch.from_numpy(mm) for mm in masks)
            model.train()
            try:
                with torch.autocast("cuda", dtype=torch.float16, enabled=mode == "fp16_autocast"):
                    logits = model(x)
            finally:
                DropPath.forced_masks = None
            tgt = torch.from_numpy(synth_targets(x.shape[0], case["num_classes"], seed=2)).cuda()
            loss = O.soft_target_ce(logits.float(), tgt, 0.1)
            (loss * scale).backward()
            sd = synth_sd(m["state_shapes"])
            for v in sd.values():
                v.requires_grad_(True)
            if is_swin:
                zo = O.swin_forward(sd, x.cpu(), patch=case["patch_size"], window=case["window_size"],
                                    depths=case["depths"], heads=case["num_heads"], drop_path_rate=case["drop_path"],
                                    training=True, masks=iter(torch.from_numpy(mm) for mm in masks))
            else:
                zo = O.vit_forward(sd, x.cpu(), patch=case["patch_size"], heads=case["num_heads"], depth=case["depth"])
            O.soft_target_ce(zo, tgt.cpu(), 0.1).backward()
            res[mode][name] = {"logits_train": rel_err(logits.detach().float(), g["logits_train" if is_swin else "logits_eval"]),
                               "worst_grad": _worst_small(model, {k: sd[k].grad for k in m["param_order"]}, scale)}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    refshim.uninstall()
    return res


def main():
    import vsn_b200  # noqa: F401
    from vsn_b200 import _lib, swin_model, vit_model
    scale = float(os.environ.get("VSN_TEST_LOSS_SCALE", "1024"))
    full = "--full" in sys.argv
    out = {"precision": _lib.PRECISION, "lib": os.path.basename(_lib.LIB_PATH), "loss_scale": scale, "cases": {}}

    for name, case in SWIN_CASES.items():
        g, m = golden(name), meta()[name]
        model = swin_model.SwinTransformerT(**swin_ctor_kwargs(case)).cuda()
        model.load_state_dict(synth_sd(m["state_shapes"], device="cuda"), strict=False)
        x = torch.from_numpy(synth_volume(case["input"], seed=1)).cuda()
        model.eval()
        with torch.no_grad():
            z = model(x)
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
        (loss * scale).backward()
        sd = synth_sd(m["state_shapes"])
        for v in sd.values():
            v.requires_grad_(True)
        zo = O.swin_forward(sd, x.cpu(), patch=case["patch_size"], window=case["window_size"], depths=case["depths"],
                            heads=case["num_heads"], drop_path_rate=case["drop_path"], training=True,
                            masks=iter(torch.from_numpy(mm) for mm in masks))
        O.soft_target_ce(zo, tgt.cpu(), 0.1).backward()
        out["cases"][name] = {"logits_eval": rel_err(z, g["logits_eval"]),
                              "logits_train": rel_err(logits.detach(), g["logits_train"]),
                              "loss": abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])),
                              "worst_grad": _worst_small(model, {k: sd[k].grad for k in m["param_order"]}, scale)}

    for name, case in VIT_CASES.items():
        g, m = golden(name), meta()[name]
        model = vit_model.ViTS(**vit_ctor_kwargs(case)).cuda()
        model.load_state_dict(synth_sd(m["state_shapes"], device="cuda"))
        x = torch.from_numpy(synth_volume(case["input"], seed=1)).cuda()
        model.train()
        logits = model(x)
        tgt = torch.from_numpy(synth_targets(x.shape[0], case["num_classes"], seed=2)).cuda()
        loss = O.soft_target_ce(logits, tgt, 0.1)
        (loss * scale).backward()
        sd = synth_sd(m["state_shapes"])
        for v in sd.values():
            v.requires_grad_(True)
        zo = O.vit_forward(sd, x.cpu(), patch=case["patch_size"], heads=case["num_heads"], depth=case["depth"])
        O.soft_target_ce(zo, tgt.cpu(), 0.1).backward()
        out["cases"][name] = {"logits_eval": rel_err(logits.detach(), g["logits_eval"]),
                              "loss": abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])),
                              "worst_grad": _worst_small(model, {k: sd[k].grad for k in m["param_order"]}, scale)}

    if full:
        g, m = golden("swin5c_full_train"), meta()["swin5c_full_train"]
        case = dict(SWIN_FULL, num_classes=5, drop_path=0.15, input=[2, 1, 144, 168, 144])
        model = swin_model.SwinTransformerT(**swin_ctor_kwargs(case)).cuda().train()
        model.load_state_dict(synth_sd(m["state_shapes"], device="cuda"), strict=False)
        x = torch.from_numpy(synth_volume(case["input"], seed=1)).cuda()
        tgt = torch.from_numpy(synth_targets(2, 5, seed=2)).cuda()
        masks = synth_keep_masks(2 * (sum(case["depths"]) - 1), 2, keep=0.7, seed=3)
        swin_model.DropPath.forced_masks = iter(torch.from_numpy(mm) for mm in masks)
        try:
            logits = model(x)
        finally:
            swin_model.DropPath.forced_masks = None
        loss = O.soft_target_ce(logits, tgt, 0.1)
        (loss * scale).backward()
        out["cases"]["swin5c_full_train"] = {"logits_train": rel_err(logits.detach(), g["logits_train"]),
                                             "loss": abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])),
                                             "worst_grad": _worst_full(model, g, scale)}
        del model
        torch.cuda.empty_cache()
        g, m = golden("vit3c_full_train"), meta()["vit3c_full_train"]
        case = dict(VIT_FULL, num_classes=3, input=[2, 1, 144, 160, 144])
        model = vit_model.ViTS(**vit_ctor_kwargs(case)).cuda().train()
        model.load_state_dict(synth_sd(m["state_shapes"], device="cuda"))
        x = torch.from_numpy(synth_volume(case["input"], seed=1)).cuda()
        tgt = torch.from_numpy(synth_targets(2, 3, seed=2)).cuda()
        logits = model(x)
        loss = O.soft_target_ce(logits, tgt, 0.1)
        (loss * scale).backward()
        out["cases"]["vit3c_full_train"] = {"logits_train": rel_err(logits.detach(), g["logits_train"]),
                                            "loss": abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])),
                                            "worst_grad": _worst_full(model, g, scale)}
    out["launches"] = _lib.launch_count()
    if "--reference" in sys.argv:
        out["reference"] = reference_modes(scale)
    print("PRECISION_CHECK " + json.dumps(out))


if __name__ == "__main__":
    main()
