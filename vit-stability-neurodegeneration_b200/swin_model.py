"""Drop-in Swin-3D modules: the reference's constructor / forward / state_dict surface
(models/swin_transformer_3d.py:106-785) on top of the vsn_b200 kernels.

The module tree only *holds* parameters (same names, shapes, dtypes, registration and initialisation
order as the reference, so `torch.manual_seed(s)` yields bit-identical initial weights and checkpoints
interchange).  `forward` never calls the sub-modules: it walks the tree and issues the fused CUDA path of
`swin.py`.  Options whose maths is not implemented by the kernels raise instead of silently differing.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import ops, swin

_VARIANTS: Dict[str, Dict[str, Union[int, List[int]]]] = {
    "T": dict(patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=[7, 7, 7]),
    "S": dict(patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=[7, 7, 7]),
    "B": dict(patch_size=[4, 4, 4], embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=[7, 7, 7]),
    "L": dict(patch_size=[4, 4, 4], embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48], window_size=[7, 7, 7]),
}


def _triple(v) -> Tuple[int, int, int]:
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)


def _unsupported(what: str):
    raise NotImplementedError(
        f"vsn_b200 Swin-3D does not implement {what}; the reference default is off "
        "(config-defaults.yaml) and there is no silent fallback")


class DropPath(nn.Module):
    """Per-sample stochastic depth (timm.layers.DropPath semantics).  Only carries `drop_prob`; the factor is
    drawn by the block and applied inside the GEMM epilogue."""

    forced_masks: Optional[Iterator[torch.Tensor]] = None   # tests inject the keep decisions here

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):  # kept for API completeness; not used on the fused path
        s = swin.droppath_scale(self.drop_prob, x.shape[0], x.device, self.training, DropPath.forced_masks)
        return x if s is None else x * s.view(-1, *([1] * (x.ndim - 1)))

    def extra_repr(self):
        return f"drop_prob={self.drop_prob:.4f}"


class MLP(nn.Sequential):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        if act_layer is not nn.GELU:
            _unsupported(f"act_layer={act_layer}")
        super().__init__(nn.Linear(in_features, hidden_features or in_features), act_layer(), nn.Dropout(drop),
                         nn.Linear(hidden_features or in_features, out_features or in_features), nn.Dropout(drop))


class WindowAttention3D(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        if dim % num_heads or dim // num_heads not in (32, 64):
            _unsupported(f"head_dim={dim / num_heads} (kernels are built for 32 and 64)")
        self.scale = (dim // num_heads) ** -0.5
        wd, wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), num_heads))
        self.register_buffer("relative_position_index", swin.relative_position_index(self.window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        self.softmax = nn.Softmax(dim=-1)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio=4.0, qkv_bias=True, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, enable_stable=False,
                 stable_lam=1.0, stable_beta=0.0, use_shakedrop=False, shakedrop_alpha_range=(-1.0, 1.0),
                 layer_scale=False, layer_scale_init_value=1e-5, post_norm=False):
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            _unsupported(f"norm_layer={norm_layer}")
        if post_norm:
            _unsupported("post_norm=True")
        if enable_stable:
            _unsupported("enable_stable=True")
        if layer_scale:
            _unsupported("layer_scale=True")
        if use_shakedrop and drop_path > 0.0:
            _unsupported("use_shakedrop=True")
        if drop > 0.0 or attn_drop > 0.0:
            _unsupported("dropout / attention_dropout > 0")
        self.dim, self.num_heads = dim, num_heads
        self.window_size, self.shift_size = tuple(window_size), tuple(shift_size)
        self.mlp_ratio, self.post_norm = mlp_ratio, post_norm
        self.enable_stable, self.stable_lam, self.stable_beta = enable_stable, stable_lam, stable_beta
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention3D(dim, self.window_size, num_heads, qkv_bias, attn_drop, drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(dim, int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.ls1 = None
        self.ls2 = None

    def gemm_weights(self) -> List[nn.Parameter]:
        return [self.attn.qkv.weight, self.attn.proj.weight, self.mlp[0].weight, self.mlp[3].weight]


class PatchEmbed3D(nn.Module):
    def __init__(self, patch_size=(4, 4, 4), in_channels=1, embed_dim=96, norm_layer=None):
        super().__init__()
        if in_channels != 1:
            _unsupported(f"in_channels={in_channels}")
        self.patch_size = tuple(patch_size)
        self.proj = nn.Conv3d(in_channels, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else nn.Identity()


class PatchMerging(nn.Module):
    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(8 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(8 * dim)


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False, **block_kw):
        super().__init__()
        self.window_size = tuple(window_size)
        self.shift_size = tuple(w // 2 for w in self.window_size)
        self.depth, self.use_checkpoint = depth, use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim=dim, num_heads=num_heads, window_size=self.window_size,
                                 shift_size=(0, 0, 0) if i % 2 == 0 else self.shift_size, mlp_ratio=mlp_ratio,
                                 qkv_bias=qkv_bias, drop=drop, attn_drop=attn_drop,
                                 drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 norm_layer=norm_layer, **block_kw)
            for i in range(depth)])
        self.downsample = downsample(dim=dim, norm_layer=norm_layer) if downsample is not None else None


class SwinTransformer3DBackbone(nn.Module):
    def __init__(self, patch_size, in_channels, embed_dim, depths, num_heads, window_size, mlp_ratio=4.0,
                 qkv_bias=True, dropout=0.0, attention_dropout=0.0, stochastic_depth_prob=0.1,
                 norm_layer=nn.LayerNorm, use_checkpoint=False, enable_stable=False, stable_k=2.0, stable_alpha=1.0,
                 use_shakedrop=False, shakedrop_alpha_range=(-1.0, 1.0), layer_scale=False,
                 layer_scale_init_value=1e-5, post_norm=False):
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            _unsupported(f"norm_layer={norm_layer}")
        self.num_layers, self.embed_dim = len(depths), embed_dim
        self.patch_size, self.window_size = _triple(patch_size), _triple(window_size)
        self.enable_stable, self.total_blocks = enable_stable, sum(depths)
        self.stable_lam, self.stable_beta = 1.0, 0.0
        self.patch_embed = PatchEmbed3D(self.patch_size, in_channels, embed_dim, norm_layer)
        self.pos_drop = nn.Dropout(p=dropout)
        dpr = [x.item() for x in torch.linspace(0, stochastic_depth_prob, sum(depths))]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i), depth=depths[i], num_heads=num_heads[i], window_size=self.window_size,
                mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=dropout, attn_drop=attention_dropout,
                drop_path=dpr[sum(depths[:i]): sum(depths[: i + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint,
                enable_stable=enable_stable, use_shakedrop=use_shakedrop,
                shakedrop_alpha_range=shakedrop_alpha_range, layer_scale=layer_scale,
                layer_scale_init_value=layer_scale_init_value, post_norm=post_norm))
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.norm = norm_layer(self.num_features)
        self.avgpool = nn.AdaptiveAvgPool3d(1)
        self.apply(self._init_weights)
        self._shadow: Optional[swin.WeightShadow] = None

    @staticmethod
    def _init_weights(m: nn.Module):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ---- fused path -----------------------------------------------------------------------
    def _gemm_params(self) -> List[nn.Parameter]:
        ps = [self.patch_embed.proj.weight]
        for layer in self.layers:
            for blk in layer.blocks:
                ps += blk.gemm_weights()
            if layer.downsample is not None:
                ps.append(layer.downsample.reduction.weight)
        return ps

    def _shadow_owner(self) -> swin.WeightShadow:
        if self._shadow is None:
            self._shadow = swin.WeightShadow(self._gemm_params())
        return self._shadow

    def forward_tokens(self, x: torch.Tensor):
        """x [B,1,D,H,W] -> (tokens fp32 [B*T, F] on the final real grid, B, T)."""
        if x.dim() != 5:
            raise ValueError(f"expected [B,1,D,H,W], got {tuple(x.shape)}")
        if type(x) is not torch.Tensor:                  # MONAI MetaTensor etc.: strip the subclass
            x = x.as_subclass(torch.Tensor)
        if not x.is_cuda:
            raise RuntimeError("vsn_b200 models run on CUDA only (there is no CPU fallback)")
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        B = x.shape[0]
        self._shadow_owner().refresh()
        wi = 0
        pe = self.patch_embed
        has_norm = isinstance(pe.norm, nn.LayerNorm)
        W = swin.GradSink.wrap
        t = swin.PatchEmbedFn.apply(x, W(pe.proj.weight), W(pe.proj.bias), W(pe.norm.weight) if has_norm else None,
                                    W(pe.norm.bias) if has_norm else None, self._shadow.view(wi), pe.patch_size)
        wi += 1
        real = tuple(-(-s // p) for s, p in zip(x.shape[2:], pe.patch_size))
        forced = DropPath.forced_masks
        # DropPath factors of the whole forward pass in two launches (one Bernoulli draw over [2 * blocks, B] with the
        # per-block keep probabilities, one division) instead of two tiny kernels per residual branch
        dp_rows = None
        if self.training and forced is None:
            probs = [blk.drop_path.drop_prob if isinstance(blk.drop_path, DropPath) else 0.0
                     for layer in self.layers for blk in layer.blocks]
            if any(q > 0.0 for q in probs):
                key = (t.device, tuple(probs))
                if getattr(self, "_dp_keep_key", None) != key:
                    keep = torch.tensor([1.0 - q for q in probs for _ in range(2)], device=t.device, dtype=torch.float32)
                    self._dp_keep, self._dp_keep_key = keep[:, None], key
                dp_rows = torch.bernoulli(self._dp_keep.expand(-1, B)) / self._dp_keep
        bi = 0
        for layer in self.layers:
            C = t.shape[1]
            pdims = swin.padded_dims(real, layer.window_size)
            if pdims != real:
                t = swin.GridCopyFn.apply(t, real, pdims, B)
            tps = pdims[0] * pdims[1] * pdims[2]
            prev_cfg = None
            for blk in layer.blocks:
                shifted = any(s > 0 for s in blk.shift_size)
                geom = ops.WindowGeom(B, pdims, blk.window_size, blk.shift_size if shifted else (0, 0, 0), shifted)
                p = blk.drop_path.drop_prob if isinstance(blk.drop_path, DropPath) else 0.0
                if dp_rows is not None:
                    s1, s2 = (dp_rows[2 * bi], dp_rows[2 * bi + 1]) if p > 0.0 else (None, None)
                else:
                    s1 = swin.droppath_scale(p, B, t.device, self.training, forced)
                    s2 = swin.droppath_scale(p, B, t.device, self.training, forced)
                bi += 1
                cfg = swin.BlockCfg(heads=blk.num_heads, hd=C // blk.num_heads, geom=geom, tokens_per_sample=tps,
                                    w16=tuple(self._shadow.view(wi + j) for j in range(4)),
                                    scale1=s1, scale2=s2)
                if prev_cfg is not None:
                    # backward hand-over of the bf16 input gradient from this block to the previous one
                    cfg.emit_for_prev, cfg.prev_scale2 = True, prev_cfg.scale2
                    prev_cfg.take_from_next = True
                prev_cfg = cfg
                wi += 4
                a = blk.attn
                t = swin.SwinBlockFn.apply(t, W(blk.norm1.weight), W(blk.norm1.bias), W(a.qkv.weight), W(a.qkv.bias),
                                           W(a.relative_position_bias_table), W(a.proj.weight), W(a.proj.bias),
                                           W(blk.norm2.weight), W(blk.norm2.bias), W(blk.mlp[0].weight),
                                           W(blk.mlp[0].bias), W(blk.mlp[3].weight), W(blk.mlp[3].bias), cfg)
            if layer.downsample is not None:
                ds = layer.downsample
                t = swin.PatchMergeFn.apply(t, W(ds.norm.weight), W(ds.norm.bias), W(ds.reduction.weight),
                                            self._shadow.view(wi), pdims, real, B, prev_cfg)
                wi += 1
                real = tuple((r + 1) // 2 for r in real)
            elif pdims != real:
                t = swin.GridCopyFn.apply(t, pdims, real, B)
        return t, B, real

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        t, B, real = self.forward_tokens(x)
        T = real[0] * real[1] * real[2]
        W = swin.GradSink.wrap
        return swin.NormPoolHeadFn.apply(t, W(self.norm.weight), W(self.norm.bias), None, None, B, T)


class SwinTransformer(nn.Module):
    def __init__(self, patch_size, in_channels, num_classes, embed_dim, depths, num_heads, window_size, mlp_ratio,
                 qkv_bias, dropout, attention_dropout, stochastic_depth_prob, norm_layer, use_checkpoint=False,
                 enable_stable=False, stable_k=2.0, stable_alpha=1.0, use_shakedrop=False,
                 shakedrop_alpha_range=(-1.0, 1.0), layer_scale=False, layer_scale_init_value=1e-5, post_norm=False):
        super().__init__()
        self.backbone = SwinTransformer3DBackbone(
            patch_size=patch_size, in_channels=in_channels, embed_dim=embed_dim, depths=depths, num_heads=num_heads,
            window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, dropout=dropout,
            attention_dropout=attention_dropout, stochastic_depth_prob=stochastic_depth_prob, norm_layer=norm_layer,
            use_checkpoint=use_checkpoint, enable_stable=enable_stable, stable_k=stable_k, stable_alpha=stable_alpha,
            use_shakedrop=use_shakedrop, shakedrop_alpha_range=shakedrop_alpha_range, layer_scale=layer_scale,
            layer_scale_init_value=layer_scale_init_value, post_norm=post_norm)
        self.head = nn.Linear(self.backbone.num_features, num_classes) if num_classes > 0 else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        bb = self.backbone
        t, B, real = bb.forward_tokens(x)
        T = real[0] * real[1] * real[2]
        W = swin.GradSink.wrap
        if isinstance(self.head, nn.Linear):
            return swin.NormPoolHeadFn.apply(t, W(bb.norm.weight), W(bb.norm.bias), W(self.head.weight),
                                             W(self.head.bias), B, T)
        return swin.NormPoolHeadFn.apply(t, W(bb.norm.weight), W(bb.norm.bias), None, None, B, T)


def _variant(name: str):
    class _V(SwinTransformer):
        def __init__(self, **kwargs):
            super().__init__(**{**_VARIANTS[name], **kwargs})
    _V.__name__ = _V.__qualname__ = f"SwinTransformer{name}"
    return _V


SwinTransformerT = _variant("T")
SwinTransformerS = _variant("S")
SwinTransformerB = _variant("B")
SwinTransformerL = _variant("L")
