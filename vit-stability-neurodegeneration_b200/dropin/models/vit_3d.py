"""Drop-in for models/vit_3d.py (reference lines 51-527): same class names and constructor kwargs."""
from vsn_b200.vit_model import *  # noqa: F401,F403
from vsn_b200 import vit_model as _m

__all__ = [n for n in dir(_m) if n.startswith("ViT") or n in ("Attention", "FeedForward", "Transformer")]
