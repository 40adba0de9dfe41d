"""Drop-in ViT-3D modules: constructor / forward / state_dict surface of models/vit_3d.py:51-527 on the
vsn_b200 kernels.  Same parameter names, shapes, registration and initialisation order as the reference."""
from __future__ import annotations

from typing import Dict, List, Literal, Optional, Tuple

import torch
import torch.nn as nn

from . import swin, vit
from .swin_model import DropPath, _unsupported

_VARIANTS: Dict[str, Dict] = {
    "S": dict(depth=12, num_heads=6, embed_dim=384, img_size=(96, 96, 96), patch_size=(16, 16, 16)),
    "B": dict(depth=12, num_heads=12, embed_dim=768, img_size=(96, 96, 96), patch_size=(16, 16, 16)),
    "L": dict(depth=24, num_heads=16, embed_dim=1024, img_size=(96, 96, 96), patch_size=(16, 16, 16)),
    "H": dict(depth=32, num_heads=16, embed_dim=1280, img_size=(96, 96, 96), patch_size=(16, 16, 16)),
}


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0, post_norm=False):
        super().__init__()
        if post_norm:
            _unsupported("post_norm=True")
        if dropout > 0.0:
            _unsupported("dropout > 0")
        self.post_norm = post_norm
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, post_norm=False):
        super().__init__()
        if post_norm:
            _unsupported("post_norm=True")
        if dropout > 0.0:
            _unsupported("attention_dropout > 0")
        if dim_head not in (32, 64):
            _unsupported(f"dim_head={dim_head} (kernels are built for 32 and 64)")
        inner = dim_head * heads
        if heads == 1 and dim_head == dim:
            _unsupported("heads == 1 and dim_head == dim (no output projection)")
        self.heads, self.dim_head, self.scale, self.post_norm = heads, dim_head, dim_head ** -0.5, post_norm
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, attention_dropout=0.0, dropout=0.0,
                 stochastic_depth_prob=0.0, use_checkpoint=False, enable_stable=False, stable_lam=1.0,
                 stable_beta=0.0, layer_scale=False, layer_scale_init_value=1e-5, post_norm=False):
        super().__init__()
        if enable_stable:
            _unsupported("enable_stable=True")
        if layer_scale:
            _unsupported("layer_scale=True")
        self.layers = nn.ModuleList([])
        self.enable_stable, self.stable_lam, self.stable_beta, self.post_norm = enable_stable, stable_lam, stable_beta, post_norm
        dpr = [x.item() for x in torch.linspace(0, stochastic_depth_prob, depth)]
        for i in range(depth):
            drop_path = DropPath(dpr[i]) if dpr[i] > 0.0 else nn.Identity()
            self.layers.append(nn.ModuleList([
                Attention(dim, heads=heads, dim_head=dim_head, dropout=attention_dropout, post_norm=post_norm),
                FeedForward(dim, mlp_dim, dropout=dropout, post_norm=post_norm),
                None, None, drop_path, None, None]))
        self.use_checkpoint = use_checkpoint


class ViT(nn.Module):
    def __init__(self, *, img_size, patch_size, num_classes, embed_dim, depth, num_heads, mlp_dim,
                 pool: Literal["cls", "mean"] = "cls", in_channels=1, dim_head=64, dropout=0.0,
                 attention_dropout=0.0, stochastic_depth_prob=0.0, use_checkpoint=False, enable_stable=False,
                 stable_k=2.0, stable_alpha=1.0, layer_scale=False, layer_scale_init_value=1e-5, post_norm=False):
        super().__init__()
        (img_d, img_h, img_w), (pd, ph, pw) = img_size, patch_size
        assert img_d % pd == 0 and img_h % ph == 0 and img_w % pw == 0, (
            f"Image dimensions ({img_d}, {img_h}, {img_w}) must be divisible by patch size ({pd}, {ph}, {pw})")
        assert pool in {"cls", "mean"}, "pool type must be either 'cls' (cls token) or 'mean' (mean pooling)"
        if in_channels != 1:
            _unsupported(f"in_channels={in_channels}")
        if enable_stable:
            _unsupported("enable_stable=True")
        num_patches = (img_d // pd) * (img_h // ph) * (img_w // pw)
        patch_dim = in_channels * pd * ph * pw
        self.patch_size = (pd, ph, pw)
        self.enable_stable, self.total_blocks, self.stable_lam, self.stable_beta = enable_stable, depth, 1.0, 0.0
        # index 0 is the parameter-free Rearrange in the reference; keep the numbering of the Sequential
        self.to_patch_embedding = nn.Sequential(nn.Identity(), nn.LayerNorm(patch_dim),
                                                nn.Linear(patch_dim, embed_dim), nn.LayerNorm(embed_dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, embed_dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.dropout = nn.Dropout(dropout)
        if dropout > 0.0:
            _unsupported("dropout > 0")
        self.transformer = Transformer(dim=embed_dim, depth=depth, heads=num_heads, dim_head=dim_head,
                                       mlp_dim=mlp_dim, attention_dropout=attention_dropout, dropout=dropout,
                                       stochastic_depth_prob=stochastic_depth_prob, use_checkpoint=use_checkpoint,
                                       enable_stable=enable_stable, layer_scale=layer_scale,
                                       layer_scale_init_value=layer_scale_init_value, post_norm=post_norm)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, num_classes))
        self.use_checkpoint = use_checkpoint
        self.apply(self._init_weights)
        self._shadow: Optional[swin.WeightShadow] = None

    @staticmethod
    def _init_weights(m: nn.Module) -> None:
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _gemm_params(self) -> List[nn.Parameter]:
        ps = [self.to_patch_embedding[2].weight]
        for layer in self.transformer.layers:
            attn, ff = layer[0], layer[1]
            ps += [attn.to_qkv.weight, attn.to_out[0].weight, ff.net[1].weight, ff.net[4].weight]
        return ps

    def _shadow_owner(self) -> swin.WeightShadow:
        if self._shadow is None:
            self._shadow = swin.WeightShadow(self._gemm_params())
        return self._shadow

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if type(x) is not torch.Tensor:
            x = x.as_subclass(torch.Tensor)
        if not x.is_cuda:
            raise RuntimeError("vsn_b200 models run on CUDA only (there is no CPU fallback)")
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        B = x.shape[0]
        for s, p in zip(x.shape[2:], self.patch_size):
            if s % p:
                raise ValueError(f"input {tuple(x.shape)} is not divisible by the patch size {self.patch_size}")
        self._shadow_owner().refresh()
        pe = self.to_patch_embedding
        W = swin.GradSink.wrap
        t = vit.ViTEmbedFn.apply(x, W(pe[1].weight), W(pe[1].bias), W(pe[2].weight), W(pe[2].bias), W(pe[3].weight),
                                 W(pe[3].bias), W(self.cls_token), W(self.pos_embedding), self._shadow.view(0),
                                 self.patch_size)
        N = t.shape[0] // B
        wi = 1
        forced = DropPath.forced_masks
        prev_cfg = None
        for layer in self.transformer.layers:
            attn, ff, dp = layer[0], layer[1], layer[4]
            p = dp.drop_prob if isinstance(dp, DropPath) else 0.0
            cfg = swin.BlockCfg(heads=attn.heads, hd=attn.dim_head, geom=None, tokens_per_sample=N, S=B, N=N,
                                w16=tuple(self._shadow.view(wi + j) for j in range(4)),
                                scale1=swin.droppath_scale(p, B, t.device, self.training, forced),
                                scale2=swin.droppath_scale(p, B, t.device, self.training, forced))
            if prev_cfg is not None:
                # backward hand-over of the bf16 input gradient from this block to the previous one (swin.BlockCfg)
                cfg.emit_for_prev, cfg.prev_scale2 = True, prev_cfg.scale2
                prev_cfg.take_from_next = True
            prev_cfg = cfg
            wi += 4
            t = swin.SwinBlockFn.apply(t, W(attn.norm.weight), W(attn.norm.bias), W(attn.to_qkv.weight), None, None,
                                       W(attn.to_out[0].weight), W(attn.to_out[0].bias), W(ff.net[0].weight),
                                       W(ff.net[0].bias), W(ff.net[1].weight), W(ff.net[1].bias), W(ff.net[4].weight),
                                       W(ff.net[4].bias), cfg)
        return vit.ViTHeadFn.apply(t, W(self.mlp_head[0].weight), W(self.mlp_head[0].bias), W(self.mlp_head[1].weight),
                                   W(self.mlp_head[1].bias), B, N, self.pool)


class ViTX(ViT):
    def __init__(self, config_name: str, img_size, patch_size, num_classes, mlp_ratio: Optional[float] = None,
                 dropout=0.0, attention_dropout=0.0, stochastic_depth_prob=0.0, in_channels=1, dim_head=64,
                 pool: Literal["cls", "mean"] = "cls", use_checkpoint=False, **overrides):
        if config_name not in _VARIANTS:
            raise ValueError(f"Unknown config_name '{config_name}'. Available: {list(_VARIANTS.keys())}")
        conf = dict(_VARIANTS[config_name])
        embed_dim = overrides.pop("embed_dim", conf["embed_dim"])
        depth = overrides.pop("depth", conf["depth"])
        num_heads = overrides.pop("num_heads", conf["num_heads"])
        ratio = mlp_ratio if mlp_ratio is not None else overrides.pop("mlp_ratio", 4.0)
        mlp_dim = overrides.pop("mlp_dim", int(embed_dim * ratio))
        super().__init__(img_size=img_size, patch_size=patch_size, num_classes=num_classes, embed_dim=embed_dim,
                         depth=depth, num_heads=num_heads, mlp_dim=mlp_dim, pool=pool, in_channels=in_channels,
                         dim_head=dim_head, dropout=dropout, attention_dropout=attention_dropout,
                         stochastic_depth_prob=stochastic_depth_prob, use_checkpoint=use_checkpoint, **overrides)


def _variant(name: str):
    class _V(ViTX):
        def __init__(self, *args, **kwargs):
            super().__init__(config_name=name, *args, **kwargs)
    _V.__name__ = _V.__qualname__ = f"ViT{name}"
    return _V


ViTS, ViTB, ViTL, ViTH = _variant("S"), _variant("B"), _variant("L"), _variant("H")
