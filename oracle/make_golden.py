#!/usr/bin/env python
"""Generate tests/golden/*.npz|json from the UNMODIFIED reference (run in the build container).

    python oracle/make_golden.py            # needs /root/reference

Imports the reference through oracle/refshim.py, loads the portable synthetic
weights/inputs of oracle/synth.py, runs forward (+backward) on CPU fp32 and
stores *outputs only*: logits, loss, per-parameter gradient norms / leading
values (full tensors for small parameters), per-stage activations' norms, the
integer known-answer hashes of SURVEY.md §4, and SAM / EMA trajectories.
The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from oracle.cases import SWIN_CASES, VIT_CASES, SWIN_FULL, VIT_FULL, swin_ctor_kwargs, vit_ctor_kwargs  # noqa: E402
from oracle.synth import synth_state, synth_volume, synth_targets, synth_keep_masks  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FULL_SMALL = 4096


def sha16(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load_synth(model, seed=0):
    sd = model.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items() if v.is_floating_point()}
    new = synth_state(shapes, seed)
    with torch.no_grad():
        for k, v in new.items():
            sd[k].copy_(torch.from_numpy(v))
    return shapes


def grads_record(model, out):
    for k, p in model.named_parameters():
        g = p.grad.detach().numpy().astype(np.float32)
        out[f"gnorm/{k}"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out[f"ghead/{k}"] = g.reshape(-1)[:8].copy()
        if g.size <= FULL_SMALL:
            out[f"gfull/{k}"] = g.copy()


def sample_index(key: str, numel: int, n: int = 256) -> np.ndarray:
    """Fixed pseudo-random element positions of a parameter (same on every box: numpy MT19937 keyed on the name)."""
    import zlib
    rs = np.random.RandomState(zlib.crc32(("gsamp/" + key).encode()) & 0xFFFFFFFF)
    return rs.randint(0, numel, size=min(n, numel)).astype(np.int64)


def run_full_size_case(name, model, x, tgt, n_dp_calls):
    """Train-mode forward + backward of the full-size model on the synthetic weights: logits, loss, and for EVERY
    parameter the gradient norm, 256 sampled values (sample_index) and the whole tensor when it has <= 1024
    elements.  This is what pins the backward of the real stage shapes (heads 3/6/12/24, 216..1 windows)."""
    from regularization.label_smoothing import LabelSmoothingLoss
    out = {}
    model.train()
    masks = synth_keep_masks(max(n_dp_calls, 1), x.shape[0], keep=0.7, seed=3)
    refshim.DropPath.forced_masks = iter([torch.from_numpy(m) for m in masks])
    model.zero_grad(set_to_none=True)
    logits = model(x)
    loss = LabelSmoothingLoss(smoothing=0.1)(logits, tgt)
    loss.backward()
    refshim.DropPath.forced_masks = None
    out["logits_train"] = logits.detach().numpy()
    out["loss"] = np.float64(loss.item())
    for k, p in model.named_parameters():
        g = p.grad.detach().numpy().astype(np.float32).reshape(-1)
        out[f"gnorm/{k}"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out[f"gsamp/{k}"] = g[sample_index(k, g.size)].copy()
        if g.size <= 1024:
            out[f"gfull/{k}"] = g.copy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "logits_train", out["logits_train"].round(5).tolist(), "loss", out["loss"])


def run_model_case(name, model, x, tgt, n_dp_calls, stage_modules, to_tokens):
    from regularization.label_smoothing import LabelSmoothingLoss
    out = {}
    taps = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, k=k: taps.__setitem__(k, to_tokens(o.detach())))
             for k, m in stage_modules.items()]
    model.eval()
    with torch.no_grad():
        out["logits_eval"] = model(x).numpy()
    for k, v in taps.items():
        out[f"tapnorm/{k}"] = np.float64(v.double().norm().item())
        out[f"taphead/{k}"] = v.reshape(-1)[:8].numpy().copy()
    for h in hooks:
        h.remove()
    model.train()
    masks = synth_keep_masks(max(n_dp_calls, 1), x.shape[0], keep=0.7, seed=3)
    refshim.DropPath.forced_masks = iter([torch.from_numpy(m) for m in masks])
    model.zero_grad(set_to_none=True)
    logits = model(x)
    loss = LabelSmoothingLoss(smoothing=0.1)(logits, tgt)
    loss.backward()
    refshim.DropPath.forced_masks = None
    out["logits_train"] = logits.detach().numpy()
    out["loss"] = np.float64(loss.item())
    grads_record(model, out)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "logits_eval", out["logits_eval"].round(5).tolist(), "loss", out["loss"])


class _Snapshot:
    def __init__(self, model):
        self.model = model

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.model.state_dict().items()}


def main():
    os.makedirs(GOLD, exist_ok=True)
    refshim.install()
    from models.swin_transformer_3d import SwinTransformerT, WindowAttention3D, BasicLayer, window_partition
    from models.vit_3d import ViTS
    from regularization.sam import SAM
    from utils.ema import EMAModel

    meta = {"torch": torch.__version__}

    # ---- integer KATs ------------------------------------------------
    win = (6, 7, 6)
    wa = WindowAttention3D(32, win, 1)
    rpi = wa.relative_position_index.numpy()
    kat = {"rpi": {"shape": list(rpi.shape), "min": int(rpi.min()), "max": int(rpi.max()), "sum": int(rpi.sum()),
                   "row0": rpi[0, :8].tolist(), "last0": int(rpi[251, 0]), "sha16": sha16(rpi)}}
    wa7 = WindowAttention3D(32, (7, 7, 7), 1).relative_position_index.numpy()
    kat["rpi_777"] = {"sha16": sha16(wa7), "sum": int(wa7.sum())}
    masks = {}
    for tag, real in (("stage0", (36, 42, 36)), ("stage1", (18, 21, 18)), ("stage2", (9, 11, 9)), ("stage3", (5, 6, 5)),
                      ("odd0", (13, 16, 12)), ("odd1", (7, 8, 6))):
        # re-run the reference's own mask code path by capturing the mask handed to block 1
        layer = BasicLayer(dim=32, depth=2, num_heads=1, window_size=win)
        cap = {}
        layer.blocks[1].register_forward_pre_hook(lambda m, a: cap.__setitem__("m", a[1]))
        with torch.no_grad():
            layer(torch.zeros(1, 32, *real))
        m = cap["m"].numpy().astype(np.float32)
        masks[tag] = {"real": list(real), "shape": list(m.shape), "nonzero": int((m != 0).sum()),
                      "windows_with_mask": int((m != 0).reshape(m.shape[0], -1).any(1).sum()),
                      "values": sorted(set(np.unique(m).tolist())), "sha16": sha16(m)}
    kat["mask"] = masks
    # window_partition order on a labelled grid
    lab = torch.arange(12 * 14 * 12, dtype=torch.float32).reshape(1, 12, 14, 12, 1)
    wp = window_partition(lab, win).squeeze(-1).numpy().astype(np.int64)
    kat["window_partition_12x14x12"] = {"sha16": sha16(wp), "w3_head": wp[3, :8].tolist()}
    meta["kat"] = kat

    # ---- small Swin cases -----------------------------------------------------------
    for name, case in SWIN_CASES.items():
        torch.manual_seed(0)
        model = SwinTransformerT(**swin_ctor_kwargs(case))
        shapes = load_synth(model)
        meta[name] = {"state_shapes": {k: list(v.shape) for k, v in model.state_dict().items()},
                      "state_dtypes": {k: str(v.dtype) for k, v in model.state_dict().items()},
                      "param_order": [k for k, _ in model.named_parameters()]}
        x = torch.from_numpy(synth_volume(case["input"], seed=1))
        tgt = torch.from_numpy(synth_targets(case["input"][0], case["num_classes"], seed=2))
        nblk = sum(case["depths"])
        n_calls = 2 * (nblk - 1) if case["drop_path"] > 0 else 0
        stages = {f"stage{i}": l for i, l in enumerate(model.backbone.layers)}
        stages["embed"] = model.backbone.patch_embed
        run_model_case(name, model, x, tgt, n_calls, stages,
                       lambda o: o.permute(0, 2, 3, 4, 1).reshape(o.shape[0], -1, o.shape[1]))

    # ---- small ViT cases ------------------------------------------------------------
    for name, case in VIT_CASES.items():
        torch.manual_seed(0)
        model = ViTS(**vit_ctor_kwargs(case))
        load_synth(model)
        meta[name] = {"state_shapes": {k: list(v.shape) for k, v in model.state_dict().items()},
                      "param_order": [k for k, _ in model.named_parameters()]}
        x = torch.from_numpy(synth_volume(case["input"], seed=1))
        tgt = torch.from_numpy(synth_targets(case["input"][0], case["num_classes"], seed=2))
        run_model_case(name, model, x, tgt, 0, {}, lambda o: o)

    # ---- full-size, reference init under manual_seed(0) ---------------
    full = {}
    torch.manual_seed(0)
    m5 = SwinTransformerT(**swin_ctor_kwargs(dict(SWIN_FULL, num_classes=5, drop_path=0.15))).eval()
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(1, 1, 144, 168, 144, generator=g).half().float()
    x2 = torch.randn(1, 1, 50, 61, 47, generator=g).half().float()
    with torch.no_grad():
        full["swin5c_seed0_in144x168x144"] = m5(x1)[0].tolist()
        full["swin5c_seed0_in50x61x47"] = m5(x2)[0].tolist()
    sd5 = m5.state_dict()
    full["swin5c_n_params"] = int(sum(p.numel() for p in m5.parameters()))
    full["swin5c_n_keys"] = len(sd5)
    full["swin5c_state_shapes"] = {k: list(v.shape) for k, v in sd5.items()}
    full["swin5c_param_sha16"] = {k: sha16(sd5[k].numpy()) for k in
                                  ("backbone.patch_embed.proj.weight", "backbone.layers.0.blocks.0.attn.qkv.weight",
                                   "backbone.layers.3.blocks.1.mlp.3.weight", "head.weight",
                                   "backbone.layers.2.blocks.5.attn.relative_position_bias_table")}
    torch.manual_seed(0)
    mv = ViTS(**vit_ctor_kwargs(dict(VIT_FULL, num_classes=3, input=None))).eval()
    g = torch.Generator().manual_seed(1)
    torch.randn(1, 1, 144, 168, 144, generator=g); torch.randn(1, 1, 50, 61, 47, generator=g)
    x3 = torch.randn(1, 1, 144, 160, 144, generator=g).half().float()
    with torch.no_grad():
        full["vit3c_seed0_in144x160x144"] = mv(x3)[0].tolist()
    sdv = mv.state_dict()
    full["vit3c_n_params"] = int(sum(p.numel() for p in mv.parameters()))
    full["vit3c_state_shapes"] = {k: list(v.shape) for k, v in sdv.items()}
    full["vit3c_param_sha16"] = {k: sha16(sdv[k].numpy()) for k in
                                 ("pos_embedding", "cls_token", "to_patch_embedding.2.weight",
                                  "transformer.layers.11.1.net.4.weight", "mlp_head.1.weight")}
    meta["full"] = full

    # ---- full-size TRAIN-mode gradients on the synthetic weights (B = 2, injected DropPath decisions) -----------
    case = dict(SWIN_FULL, num_classes=5, drop_path=0.15, input=[2, 1, 144, 168, 144])
    torch.manual_seed(0)
    model = SwinTransformerT(**swin_ctor_kwargs(case))
    load_synth(model)
    x = torch.from_numpy(synth_volume(case["input"], seed=1))
    tgt = torch.from_numpy(synth_targets(2, 5, seed=2))
    run_full_size_case("swin5c_full_train", model, x, tgt, 2 * (sum(case["depths"]) - 1))
    meta["swin5c_full_train"] = {"state_shapes": {k: list(v.shape) for k, v in model.state_dict().items()},
                                 "param_order": [k for k, _ in model.named_parameters()]}
    del model
    vcase = dict(VIT_FULL, num_classes=3, input=[2, 1, 144, 160, 144])
    torch.manual_seed(0)
    model = ViTS(**vit_ctor_kwargs(vcase))
    load_synth(model)
    x = torch.from_numpy(synth_volume(vcase["input"], seed=1))
    tgt = torch.from_numpy(synth_targets(2, 3, seed=2))
    run_full_size_case("vit3c_full_train", model, x, tgt, 0)
    meta["vit3c_full_train"] = {"state_shapes": {k: list(v.shape) for k, v in model.state_dict().items()},
                                "param_order": [k for k, _ in model.named_parameters()]}
    del model
    print("full", {k: v for k, v in full.items() if "seed0" in k})

    # ---- SAM / EMA trajectories on a tiny parameter set --------------------------------
    rs = np.random.RandomState(7)
    shapes = [(5, 3), (7,), (2, 3, 4), (1,)]
    p0 = [rs.standard_normal(s).astype(np.float32) for s in shapes]
    g0 = [rs.standard_normal(s).astype(np.float32) for s in shapes]
    sam_out = {}
    for adaptive in (False, True):
        ps = [torch.nn.Parameter(torch.from_numpy(p.copy())) for p in p0]
        opt = SAM([{"params": ps[:2]}, {"params": ps[2:], "weight_decay": 0.0}], torch.optim.AdamW,
                  rho=0.05, adaptive=adaptive, lr=1e-3, weight_decay=0.05)
        for p, g in zip(ps, g0):
            p.grad = torch.from_numpy(g.copy())
        sam_out[f"norm_{int(adaptive)}"] = np.float64(opt._grad_norm().item())
        opt.first_step(zero_grad=True)
        for i, p in enumerate(ps):
            sam_out[f"pert_{int(adaptive)}_{i}"] = p.detach().numpy().copy()
        for p, g in zip(ps, g0):
            p.grad = torch.from_numpy((0.5 * g).copy())
        opt.second_step(zero_grad=True)
        for i, p in enumerate(ps):
            sam_out[f"final_{int(adaptive)}_{i}"] = p.detach().numpy().copy()
    for i, (p, g) in enumerate(zip(p0, g0)):
        sam_out[f"p0_{i}"], sam_out[f"g0_{i}"] = p, g
    # non-finite and zero gradients (regularization/sam.py:46-52,66-70,141-153): "inf1" = one tensor holds an inf,
    # "nan2" = another holds a nan, "allbad" = every tensor is non-finite, "zero" = all gradients are zero
    def bad_grads(tag):
        gs = [g.copy() for g in g0]
        if tag == "inf1":
            gs[1][3] = np.inf
        elif tag == "nan2":
            gs[2][0, 1, 2] = np.nan
        elif tag == "allbad":
            for g in gs:
                g.reshape(-1)[0] = np.inf
        elif tag == "zero":
            gs = [np.zeros_like(g) for g in gs]
        return gs
    for tag in ("inf1", "nan2", "allbad", "zero"):
        ps = [torch.nn.Parameter(torch.from_numpy(p.copy())) for p in p0]
        opt = SAM([{"params": ps}], torch.optim.AdamW, rho=0.05, adaptive=False, lr=1e-3, weight_decay=0.05)
        for p, g in zip(ps, bad_grads(tag)):
            p.grad = torch.from_numpy(g)
        sam_out[f"nf_{tag}_norm"] = np.float64(opt._grad_norm().item())
        opt.first_step(zero_grad=False)
        for i, p in enumerate(ps):
            sam_out[f"nf_{tag}_pert_{i}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "sam.npz"), **sam_out)

    lin = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
    with torch.no_grad():
        for p in lin.parameters():
            p.copy_(torch.from_numpy(rs.standard_normal(tuple(p.shape)).astype(np.float32)))
    ema = EMAModel(lin, decay=0.999)
    ema_out = {}
    for k, v in lin.state_dict().items():
        ema_out[f"s0/{k}"] = v.numpy().copy()
    for step in range(1, 6):
        with torch.no_grad():
            for p in lin.parameters():
                p.add_(torch.from_numpy(rs.standard_normal(tuple(p.shape)).astype(np.float32)) * 0.1)
            lin[1].num_batches_tracked += 1
        # On CUDA the reference snapshots with `.cpu()` (a copy, utils/ema.py:84-87); on a CPU model
        # `v.detach()` would alias the live parameters.  Hand it copies so the golden reflects the
        # CUDA semantics the trainer actually runs.
        ema.update(_Snapshot(lin))
        for k, v in lin.state_dict().items():
            ema_out[f"s{step}/{k}"] = v.numpy().copy()
        for k, v in ema.model_state.items():
            ema_out[f"ema{step}/{k}"] = v.numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "ema.npz"), **ema_out)

    with open(os.path.join(GOLD, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(GOLD)))


if __name__ == "__main__":
    main()
