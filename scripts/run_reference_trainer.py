#!/usr/bin/env python
"""Drive the UNMODIFIED reference trainer (train/train_transformer.py) through vsn_b200's drop-in packages on
synthetic data: proof that `train/train_transformer.py`, `regularization/sam.py` users and
`utils/ema.py` users run unchanged on the sm_100a path.  TEST / MEASUREMENT HARNESS, not product.

    python scripts/run_reference_trainer.py --arch swin --steps 6 --out gpurun_out/trainer_swin
    torchrun --nproc-per-node 2 ... scripts/run_reference_trainer.py --arch swin --steps 6 --out ...

What it sets up, all outside the reference's files:
  * `monai` / `nibabel` / `nilearn` / `timm` stand-ins (absent from this image): Compose, NormalizeIntensity and Resize
    with MONAI's semantics for the minimal-augmentation branch (train/train_transformer.py:1729-1752), no-op Rand*;
  * a synthetic cohort: 10 fold CSVs (Subject, Diagnosis) + fp16 `[1,D,H,W]` `.pt` volumes in the cache directory, so
    the trainer's preprocessing step finds everything present (train/train_transformer.py:1571-1576);
  * a config YAML overriding the chosen `configs/*-no_seed-baseline.yaml`: few steps, small volumes, SAM + EMA +
    MixUp + balanced sampler on, validation every 2 steps, checkpoints kept;
  * `sys.path`: vsn_b200's `dropin/` first, the reference root last (as the trainer itself appends it), CWD = the
    reference root (wandb loads `config-defaults.yaml` from there).
Then `runpy` executes the trainer as `__main__`.  With `--resume` the run is repeated from the `_last.pt` checkpoint
the first run wrote; `--compare` loads that checkpoint into the reference's own model class on the CPU and compares
its logits with the drop-in model's on the same input.

`--eval` drives the UNMODIFIED evaluation script (eval/eval_transformer.py: build_model from the run's saved W&B config,
load_checkpoint, channels_last_3d inputs under torch.inference_mode, optional fp16 autocast, bootstrap metrics, the
prediction CSVs) over the checkpoints the trainer wrote, the drop-in modules shadowing the reference's;
`--eval --reference-models` is the control run (the reference's own modules, any device), and `--eval-compare` checks
the two runs' prediction CSVs against each other (same checkpoints, same subjects: class probabilities).
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def install_stubs():
    import torch
    from oracle import refshim
    ref_root = refshim.install()            # timm.layers + minimal monai (Transform, MapTransform, set_determinism)
    sys.path.remove(ref_root)
    mt = sys.modules["monai.transforms"]

    class Compose:
        def __init__(self, transforms=None, **kw):
            self.transforms = list(transforms or [])

        def __call__(self, x):
            for t in self.transforms:
                x = t(x)
            return x

    class NormalizeIntensity:                # monai default: (x - mean) / std over the whole image, std == 0 -> no division
        def __init__(self, *a, **kw):
            pass

        def __call__(self, x):
            x = x.float()
            s = x.std(unbiased=False)
            return (x - x.mean()) / s if float(s) != 0.0 else x - x.mean()

    class Resize:                            # monai default mode "area" on [C, D, H, W]
        def __init__(self, spatial_size, *a, **kw):
            self.size = tuple(int(v) for v in spatial_size)

        def __call__(self, x):
            if tuple(x.shape[-3:]) == self.size:
                return x
            return torch.nn.functional.interpolate(x[None].float(), size=self.size, mode="area")[0]

    class _Identity:
        def __init__(self, *a, **kw):
            pass

        def __call__(self, x):
            return x

    mt.Compose, mt.NormalizeIntensity, mt.Resize = Compose, NormalizeIntensity, Resize
    for name in ("Rand3DElastic", "RandAdjustContrast", "RandAffine", "RandBiasField", "RandFlip", "RandGibbsNoise",
                 "RandHistogramShift", "RandKSpaceSpikeNoise", "RandScaleIntensity", "CenterSpatialCrop", "OneOf",
                 "Identity", "RandSpatialCrop", "Flip", "Affine"):
        setattr(mt, name, type(name, (_Identity,), {}))
    for name in ("nibabel", "nilearn", "nilearn.image", "nilearn.masking"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["nilearn"].image = sys.modules["nilearn.image"]
    sys.modules["nilearn"].masking = sys.modules["nilearn.masking"]

    def _unavailable(*a, **kw):
        raise RuntimeError("NIfTI preprocessing is not part of the harness (the .pt cache is synthetic)")
    sys.modules["nilearn.image"].resample_img = _unavailable
    sys.modules["nibabel"].load = _unavailable
    return ref_root


def make_cohort(out, img_size, classes, per_fold=4):
    """10 fold CSVs + fp16 [1, D, H, W] volumes (dataset/preprocessing.py:242-249 cache format)."""
    import numpy as np
    import pandas as pd
    import torch
    csv_dir, cache = os.path.join(out, "folds"), os.path.join(out, "cache", "train")
    os.makedirs(csv_dir, exist_ok=True)
    os.makedirs(cache, exist_ok=True)
    rs = np.random.RandomState(0)
    n = 0
    for f in range(10):
        rows = []
        for i in range(per_fold):
            subj = f"sub{f:02d}{i:02d}"
            cls = classes[(f + i) % len(classes)]
            rows.append({"Subject": subj, "Diagnosis": cls, "Dataset": "SYNTH", "Age": 60 + i, "Sex": "F"})
            p = os.path.join(cache, subj + ".pt")
            if not os.path.exists(p):
                v = rs.standard_normal((1, *img_size)).astype(np.float32) + 0.3 * classes.index(cls)
                torch.save(torch.from_numpy(v).half(), p)
            n += 1
        pd.DataFrame(rows).to_csv(os.path.join(csv_dir, f"fold_{f}.csv"), index=False)
    return csv_dir, os.path.join(out, "cache"), n


def write_config(out, ref_root, arch, img_size, steps, classes):
    import yaml
    base = {"swin": f"swin-{len(classes)}c-no_seed-baseline.yaml", "vit": f"vit-{len(classes)}c-no_seed-baseline.yaml"}[arch]
    with open(os.path.join(ref_root, "configs", base)) as f:
        cfg = yaml.safe_load(f)
    over = {"IMG_SIZE": list(img_size), "RESHAPE_SIZE": list(img_size) if arch == "vit" else None, "STEPS": steps,
            "BATCH_SIZE": 2, "EFFECTIVE_BATCH_SIZE": 8, "USE_EMA": True, "USE_SAM": True, "USE_MIXUP": True,
            "USE_BALANCED_SAMPLER": True, "VALIDATION_FREQUENCY": 2, "NUM_WORKERS": 0, "PREFETCH_FACTOR": None,
            "PRELOAD_DATA": True, "KEEP_BEST_N": 2, "LR_WARMUP": 1, "WD_WARMUP": 1, "EARLY_STOPPING_PATIENCE": 1000,
            "SEED": 123}
    for k, v in over.items():
        if k in cfg and isinstance(cfg[k], dict) and "value" in cfg[k]:
            cfg[k]["value"] = v
        else:
            cfg[k] = {"value": v}
    path = os.path.join(out, f"{arch}-harness.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path


def compare_checkpoint(out, arch, img):
    """Load the checkpoint the trainer wrote (its "model" entry is the EMA state, train/train_transformer.py:808) into
    the REFERENCE's own model class on the CPU and into the vsn_b200 drop-in on the GPU; same input, compare logits."""
    import glob
    import torch
    ck = sorted(glob.glob(os.path.join(out, "runs", f"{arch}_harness", "*_last.pt")))
    if not ck:
        raise SystemExit("compare: no *_last.pt checkpoint")
    ckpt = torch.load(ck[-1], map_location="cpu", weights_only=False)
    sd = ckpt["model"]
    from oracle import refshim
    refshim.install()
    if arch == "swin":
        from models.swin_transformer_3d import SwinTransformerT as RefModel
        kw = dict(in_channels=1, patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                  window_size=[6, 7, 6], mlp_ratio=4.0, qkv_bias=True, dropout=0.0, attention_dropout=0.0,
                  stochastic_depth_prob=0.15, num_classes=3, norm_layer=torch.nn.LayerNorm)
    else:
        from models.vit_3d import ViTS as RefModel
        kw = dict(img_size=tuple(img), num_classes=3, in_channels=1, patch_size=(16, 16, 16), mlp_ratio=4.0, dropout=0.0,
                  attention_dropout=0.0, embed_dim=384, num_heads=6, depth=12)
    ref = RefModel(**kw).eval()
    missing, unexpected = ref.load_state_dict(sd, strict=True), None
    refshim.uninstall()
    import vsn_b200  # noqa: F401
    from vsn_b200 import swin_model, vit_model
    Ours = swin_model.SwinTransformerT if arch == "swin" else vit_model.ViTS
    ours = Ours(**kw).cuda().eval()
    ours.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 1, *img, generator=g).half().float()
    with torch.no_grad():
        zr = ref(x)
        zo = ours(x.cuda()).cpu()
    err = float((zo - zr).norm() / zr.norm())
    print(f"compare: checkpoint {os.path.basename(ck[-1])} (step {ckpt.get('step')}), {len(sd)} tensors loaded strictly into "
          f"the reference {RefModel.__name__} (CPU fp32) and the vsn_b200 drop-in (GPU): logits rel err {err:.3e}")
    print("compare: reference logits", [round(float(v), 4) for v in zr[0]], "drop-in", [round(float(v), 4) for v in zo[0]])
    assert err < 2e-2, err


def run_eval(a, out, ref_root):
    """eval/eval_transformer.py as __main__ over every checkpoint of the harness run (its own argv, its own W&B config
    lookup next to the checkpoints: train/train_transformer.py passes dir=save_dir to wandb.init)."""
    import glob
    save_dir = os.path.join(out, "runs", f"{a.arch}_harness")
    cks = sorted(glob.glob(os.path.join(save_dir, "model_*_last.pt")) + glob.glob(os.path.join(save_dir, "model_*_best*.pt")))
    if a.eval_last_only:
        cks = [c for c in cks if c.endswith("_last.pt")]
    if not cks:
        raise SystemExit("eval: the trainer harness has written no checkpoint under " + save_dir)
    if not a.reference_models:
        sys.path.insert(0, os.path.join(ROOT, "vit-stability-neurodegeneration_b200", "dropin"))
    sys.path.append(ref_root)
    folder = a.eval_folder or ("eval_reference" if a.reference_models else "eval_dropin")
    argv = ["eval_transformer.py", "--training-csv-dir", os.path.join(out, "folds"), "--intermediate-dir",
            os.path.join(out, "cache"), "--checkpoints", *cks, "--batch-size", "2", "--force-eval", "--output-folder", folder]
    if a.use_amp:
        argv.append("--use-amp")
    os.chdir(ref_root)
    os.environ.setdefault("WANDB_SILENT", "true")
    # utils/bootstrap_metric.py:594-598 picks joblib's threading backend on its cluster (SLURM_JOB_ID set) and loky
    # elsewhere; loky's worker processes cannot import the harness's monai / timm stand-ins, so the harness takes the
    # reference's own cluster branch
    os.environ.setdefault("SLURM_JOB_ID", "harness")
    # ... and the metric post-processing (10000 bootstrap resamples per split, minutes under the GIL) is cut to 200: the
    # script's `from utils import compute_bootstrap_metrics` binds this wrapper; nothing on the model path changes
    import functools
    import utils as ref_utils
    ref_utils.compute_bootstrap_metrics = functools.partial(ref_utils.compute_bootstrap_metrics, n_bootstrap=a.bootstrap)
    sys.argv = argv
    print(f"harness: eval_transformer.py over {len(cks)} checkpoints -> {os.path.join(save_dir, folder)}", flush=True)
    runpy.run_path(os.path.join(ref_root, "eval", "eval_transformer.py"), run_name="__main__")


def compare_eval(out, arch, tol):
    """The prediction CSVs eval_transformer.py wrote through the drop-in against the control run's (reference modules)."""
    import glob
    import numpy as np
    import pandas as pd
    save_dir = os.path.join(out, "runs", f"{arch}_harness")
    ours = sorted(glob.glob(os.path.join(save_dir, "eval_dropin", "prediction_*_id.csv")))
    if not ours:
        raise SystemExit("eval-compare: no drop-in predictions")
    worst = 0.0
    for p in ours:
        q = os.path.join(save_dir, "eval_reference", os.path.basename(p))
        if not os.path.exists(q):
            raise SystemExit("eval-compare: no control prediction file " + q)
        da, db = pd.read_csv(p), pd.read_csv(q)
        assert list(da["Subject"]) == list(db["Subject"]), "subjects differ"
        cols = [c for c in da.columns if c not in db.columns or da[c].dtype.kind == "f"]
        cols = [c for c in cols if c in db.columns and db[c].dtype.kind == "f" and c not in ("Age",)]
        pa, pb = da[cols].to_numpy(dtype=np.float64), db[cols].to_numpy(dtype=np.float64)
        err = float(np.abs(pa - pb).max())
        worst = max(worst, err)
        print(f"eval-compare: {os.path.basename(p)}: {len(da)} subjects, columns {cols}, max |dp| {err:.3e}; "
              f"same argmax {int((pa.argmax(1) == pb.argmax(1)).sum())}/{len(da)}")
    assert worst < tol, (worst, tol)
    print(f"eval-compare: ok (worst {worst:.3e} < {tol})")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="swin", choices=["swin", "vit"])
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trainer_harness"))
    ap.add_argument("--img-size", type=int, nargs=3, default=None)
    ap.add_argument("--resume", action="store_true", help="resume from the _last.pt checkpoint of a previous run")
    ap.add_argument("--compare", action="store_true",
                    help="no training: load the run's _last.pt into the reference model (CPU) and the drop-in (GPU)")
    ap.add_argument("--eval", action="store_true", help="run the unmodified eval/eval_transformer.py over the run's checkpoints")
    ap.add_argument("--eval-compare", action="store_true", help="compare the drop-in and the control prediction CSVs")
    ap.add_argument("--eval-last-only", action="store_true", help="eval: only the *_last.pt checkpoint")
    ap.add_argument("--eval-folder", default=None)
    ap.add_argument("--bootstrap", type=int, default=200, help="eval: bootstrap resamples of the metric post-processing")
    ap.add_argument("--eval-tol", type=float, default=2e-2, help="largest allowed |difference| of a class probability")
    ap.add_argument("--use-amp", action="store_true", help="eval: pass --use-amp (fp16 autocast around the model call)")
    ap.add_argument("--reference-models", action="store_true",
                    help="do NOT shadow the reference's modules (control run of the harness itself, any device)")
    a = ap.parse_args()
    out = os.path.abspath(a.out)
    os.makedirs(out, exist_ok=True)
    classes = ["CN", "AD", "FTD"]
    img = tuple(a.img_size) if a.img_size else ((48, 56, 48) if a.arch == "swin" else (48, 64, 48))
    if a.compare:
        return compare_checkpoint(out, a.arch, img)
    if a.eval_compare:
        return compare_eval(out, a.arch, a.eval_tol)
    ref_root = install_stubs()
    if a.eval:
        return run_eval(a, out, ref_root)
    rank = int(os.environ.get("RANK", "0"))
    if rank == 0:
        csv_dir, cache, n = make_cohort(out, img, classes)
        print(f"harness: {n} synthetic subjects {img} in {cache}", flush=True)
    else:
        csv_dir, cache = os.path.join(out, "folds"), os.path.join(out, "cache")
    cfg = write_config(out, ref_root, a.arch, img, a.steps, classes) if rank == 0 else os.path.join(out, f"{a.arch}-harness.yaml")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:      # the other ranks wait for the cohort
        import time
        while not os.path.exists(cfg):
            time.sleep(0.2)
        time.sleep(1.0)
    if not a.reference_models:
        sys.path.insert(0, os.path.join(ROOT, "vit-stability-neurodegeneration_b200", "dropin"))
    sys.path.append(ref_root)                            # what train_transformer.py:51 does itself
    save_dir = os.path.join(out, "runs")
    argv = ["train_transformer.py", "--training-csv-dir", csv_dir, "--intermediate-dir", cache, "--save-dir", save_dir,
            "--runname", f"{a.arch}_harness", "--wandb-mode", "offline", "--config", cfg, "--fold", "0"]
    if a.resume:
        import glob
        last = sorted(glob.glob(os.path.join(save_dir, f"{a.arch}_harness", "*_last.pt")))
        if not last:
            raise SystemExit("no *_last.pt checkpoint to resume from")
        argv += ["--checkpoint", last[-1]]
    os.chdir(ref_root)                                   # wandb reads ./config-defaults.yaml
    os.environ.setdefault("WANDB_SILENT", "true")
    os.environ.setdefault("WANDB_DIR", out)
    sys.argv = argv
    runpy.run_path(os.path.join(ref_root, "train", "train_transformer.py"), run_name="__main__")


if __name__ == "__main__":
    main()
