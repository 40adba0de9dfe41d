"""Drop-in for models/swin_transformer_3d.py (reference lines 106-785): same class names and constructor kwargs."""
from vsn_b200.swin_model import (BasicLayer, DropPath, MLP, PatchEmbed3D, PatchMerging, SwinTransformer,  # noqa: F401
                                 SwinTransformer3DBackbone, SwinTransformerB, SwinTransformerBlock, SwinTransformerL,
                                 SwinTransformerS, SwinTransformerT, WindowAttention3D)
