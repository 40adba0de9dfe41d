"""Import shim for the UNMODIFIED reference (test infrastructure, not product).

The reference (`/root/reference`, or a copy under `baseline/_ref/`) imports
`timm.layers` and `monai`, which are not installed here.  SURVEY.md §8c /
Appendix B lists the exact semantics of the three timm helpers on the model
path; this module registers stand-ins with those semantics and puts the
reference root on `sys.path` so `models.swin_transformer_3d`, `models.vit_3d`,
`regularization.sam` and `utils.ema` import as they are.

Only `oracle/make_golden.py`, `tests/` and `bench.py --impl reference` use it.
"""
from __future__ import annotations

import os
import sys
import types

import torch


class _ShimDropPath(torch.nn.Module):
    """timm.layers.DropPath semantics: per-sample Bernoulli keep
    mask scaled by 1/keep in training, identity otherwise.  `forced_masks` lets
    a test inject the keep decisions so two implementations see the same draw."""

    forced_masks = None  # optional iterator of [B] float tensors

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        if _ShimDropPath.forced_masks is not None:
            m = next(_ShimDropPath.forced_masks).to(x).view(shape)
        else:
            m = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            m = m / keep
        return x * m


def _to_3tuple(x):
    if isinstance(x, (list, tuple)):
        return tuple(x)
    return (x, x, x)


def reference_root() -> str | None:
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("VSN_REFERENCE_ROOT"), os.path.join(here, "..", "baseline", "_ref"),
                 "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "models")):
            return os.path.abspath(cand)
    return None


def install() -> str:
    """Register the stubs and make the reference importable.  Returns its root."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found (neither /root/reference nor baseline/_ref)")
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        layers = types.ModuleType("timm.layers")
        layers.DropPath = _ShimDropPath
        layers.to_3tuple = _to_3tuple
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.layers"] = layers
    if "monai" not in sys.modules:
        monai = types.ModuleType("monai")
        mu = types.ModuleType("monai.utils")
        mu.set_determinism = lambda seed=None, **kw: None
        mt = types.ModuleType("monai.transforms")
        mt.Transform = type("Transform", (), {})
        mt.MapTransform = type("MapTransform", (), {})
        monai.utils, monai.transforms = mu, mt
        sys.modules.update({"monai": monai, "monai.utils": mu, "monai.transforms": mt})
    # The reference's package names (models, utils, regularization) are generic;
    # purge any same-named modules loaded from elsewhere, then put it first.
    for name in list(sys.modules):
        top = name.split(".")[0]
        if top in ("models", "utils", "regularization", "dataset"):
            f = getattr(sys.modules[name], "__file__", "") or ""
            if not f.startswith(root):
                del sys.modules[name]
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    return root


def uninstall() -> None:
    root = reference_root()
    if root and root in sys.path:
        sys.path.remove(root)
    for name in list(sys.modules):
        top = name.split(".")[0]
        if top in ("models", "utils", "regularization", "dataset"):
            del sys.modules[name]


DropPath = _ShimDropPath
