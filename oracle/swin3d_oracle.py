"""CPU oracle for the Swin-3D / ViT-3D training hot path.  TEST INFRASTRUCTURE ONLY.

A functional restatement (plain functions over a ``state_dict``-shaped mapping,
CPU torch + numpy, fp32 or fp64) of what the reference computes on the path
named in SURVEY.md §8a.  It is deliberately *not* structured like the
reference: windows are gathered with explicit token-index tables, the shift
mask and the relative-position index come from closed forms, and there are no
``nn.Module``s.  torch is used only as an array library with autograd so the
same function yields the gradients.

Pinned against goldens generated from the UNMODIFIED reference by
``oracle/make_golden.py`` (fixtures in ``tests/golden/``); the reference itself
ships no tests or golden vectors for this path, so at the
reference-test level parity is "unpinned" and these fixtures are the pin.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product never does.

Reference lines followed (all under /root/reference):
  models/swin_transformer_3d.py:52-69    MLP
  models/swin_transformer_3d.py:72-89    window_partition / window_reverse
  models/swin_transformer_3d.py:132-152  relative_position_index
  models/swin_transformer_3d.py:162-199  WindowAttention3D.forward
  models/swin_transformer_3d.py:328-380  SwinTransformerBlock.forward (pre-norm)
  models/swin_transformer_3d.py:453-514  BasicLayer.forward (pad, mask, crop)
  models/swin_transformer_3d.py:532-543  PatchEmbed3D.forward
  models/swin_transformer_3d.py:553-572  PatchMerging.forward
  models/swin_transformer_3d.py:685-698,758-761  backbone tail + head
  models/vit_3d.py:110-142,204-255,364-374,438-457  ViT
  regularization/sam.py:38-155           SAM
  utils/ema.py:72-108                    EMA (last-3 weighted average)
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# integer index machinery (bit-exact part of the contract)
# ----------------------------------------------------------------------------

def relative_position_index(window: Sequence[int]) -> np.ndarray:
    """int64 [N,N] with N = wd*wh*ww.  Closed form of
    models/swin_transformer_3d.py:132-152:
    idx[i,j] = (di-dj+wd-1)*(2wh-1)(2ww-1) + (hi-hj+wh-1)*(2ww-1) + (wi-wj+ww-1)."""
    wd, wh, ww = window
    n = wd * wh * ww
    t = np.arange(n)
    d, h, w = t // (wh * ww), (t // ww) % wh, t % ww
    lin = d * ((2 * wh - 1) * (2 * ww - 1)) + h * (2 * ww - 1) + w
    off = (wd - 1) * (2 * wh - 1) * (2 * ww - 1) + (wh - 1) * (2 * ww - 1) + (ww - 1)
    return (lin[:, None] - lin[None, :] + off).astype(np.int64)


def _axis_region(size: int, win: int, shift: int) -> np.ndarray:
    """Region label 0/1/2 along one axis of the padded grid:
    [0,size-win) | [size-win,size-shift) | [size-shift,size)
    (models/swin_transformer_3d.py:466-480)."""
    r = np.zeros(size, dtype=np.int64)
    r[size - win: size - shift] = 1
    r[size - shift:] = 2
    return r


def window_tokens(grid: Sequence[int], window: Sequence[int]) -> np.ndarray:
    """int64 [nW, N]: flat index (d*Hp+h)*Wp+w of token n of window k on the grid,
    in the order window_partition produces (models/swin_transformer_3d.py:72-79)."""
    Dp, Hp, Wp = grid
    wd, wh, ww = window
    nd, nh, nw = Dp // wd, Hp // wh, Wp // ww
    out = np.empty((nd * nh * nw, wd * wh * ww), dtype=np.int64)
    n = np.arange(wd * wh * ww)
    ld, lh, lw = n // (wh * ww), (n // ww) % wh, n % ww
    k = 0
    for a in range(nd):
        for b in range(nh):
            for c in range(nw):
                out[k] = ((a * wd + ld) * Hp + (b * wh + lh)) * Wp + (c * ww + lw)
                k += 1
    return out


def shifted_source(grid: Sequence[int], shift: Sequence[int]) -> np.ndarray:
    """int64 [Dp*Hp*Wp]: rolled[p] = original[src[p]] for torch.roll by -shift
    (models/swin_transformer_3d.py:333-341): src(d,h,w) = ((d+sd)%Dp, ...)."""
    Dp, Hp, Wp = grid
    d = (np.arange(Dp) + shift[0]) % Dp
    h = (np.arange(Hp) + shift[1]) % Hp
    w = (np.arange(Wp) + shift[2]) % Wp
    return ((d[:, None, None] * Hp + h[None, :, None]) * Wp + w[None, None, :]).reshape(-1)


def region_ids(grid: Sequence[int], window: Sequence[int], shift: Sequence[int]) -> np.ndarray:
    """int64 [Dp*Hp*Wp] region label 0..26 of every position of the ROLLED grid."""
    Dp, Hp, Wp = grid
    rd = _axis_region(Dp, window[0], shift[0])
    rh = _axis_region(Hp, window[1], shift[1])
    rw = _axis_region(Wp, window[2], shift[2])
    return (9 * rd[:, None, None] + 3 * rh[None, :, None] + rw[None, None, :]).reshape(-1)


def shift_mask(grid: Sequence[int], window: Sequence[int], shift: Sequence[int]) -> np.ndarray:
    """float32 [nW,N,N] with exactly {0,-100} (models/swin_transformer_3d.py:463-492)."""
    reg = region_ids(grid, window, shift)[window_tokens(grid, window)]  # [nW,N]
    return np.where(reg[:, :, None] != reg[:, None, :], np.float32(-100.0), np.float32(0.0))


def padded_grid(real: Sequence[int], window: Sequence[int]) -> Tuple[int, int, int]:
    return tuple(int(math.ceil(r / w) * w) for r, w in zip(real, window))


# ----------------------------------------------------------------------------
# floating-point pieces
# ----------------------------------------------------------------------------

def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _gelu(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _droppath(z: Tensor, masks: Optional[Iterator[Tensor]], p: float, training: bool) -> Tensor:
    """timm DropPath semantics.  `masks` yields the [B] keep
    decisions in call order; p==0 or eval ⇒ identity and no draw is consumed."""
    if not training or p == 0.0:
        return z
    m = next(masks).to(z).view(-1, *([1] * (z.ndim - 1)))
    return z * (m / (1.0 - p))


def window_attention(y: Tensor, sd: Dict[str, Tensor], pre: str, heads: int,
                     grid, window, shift, shifted: bool) -> Tensor:
    """y: [B, Dp*Hp*Wp, C] (already LayerNorm-ed).  Returns the attention branch
    output in the same (un-rolled, un-windowed) token order."""
    B, T, C = y.shape
    hd = C // heads
    tok = torch.from_numpy(window_tokens(grid, window))          # [nW,N] positions on rolled grid
    nW, N = tok.shape
    if shifted:
        src = torch.from_numpy(shifted_source(grid, shift))[tok]  # positions on original grid
    else:
        src = tok
    xw = y[:, src.reshape(-1), :].reshape(B * nW, N, C)
    qkv = xw @ sd[pre + "qkv.weight"].t()
    if (pre + "qkv.bias") in sd:
        qkv = qkv + sd[pre + "qkv.bias"]
    qkv = qkv.reshape(B * nW, N, 3, heads, hd)
    q = qkv[:, :, 0].permute(0, 2, 1, 3) * (hd ** -0.5)
    k = qkv[:, :, 1].permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)
    s = q @ k.transpose(-1, -2)                                   # [B*nW, h, N, N]
    rpi = torch.from_numpy(relative_position_index(window))
    bias = sd[pre + "relative_position_bias_table"][rpi.reshape(-1)].reshape(N, N, heads)
    s = s + bias.permute(2, 0, 1)[None]
    if shifted:
        m = torch.from_numpy(shift_mask(grid, window, shift)).to(s.dtype)   # [nW,N,N]
        s = (s.reshape(B, nW, heads, N, N) + m[None, :, None]).reshape(B * nW, heads, N, N)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * nW, N, C)
    o = o @ sd[pre + "proj.weight"].t() + sd[pre + "proj.bias"]
    out = torch.zeros_like(y)
    out[:, src.reshape(-1), :] = o.reshape(B, nW * N, C)          # window_reverse + roll back
    return out


def swin_forward(sd: Dict[str, Tensor], x: Tensor, *, patch=(4, 4, 4), window=(6, 7, 6),
                 depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), drop_path_rate: float = 0.0,
                 training: bool = False, masks: Optional[Iterator[Tensor]] = None,
                 taps: Optional[dict] = None) -> Tensor:
    """x: [B,1,D,H,W] → logits [B,K].  `sd` uses the reference's state_dict keys."""
    B = x.shape[0]
    pd, ph, pw = patch
    D, H, W = x.shape[2:]
    x = F.pad(x, (0, (-W) % pw, 0, (-H) % ph, 0, (-D) % pd))
    D, H, W = x.shape[2] // pd, x.shape[3] // ph, x.shape[4] // pw
    # patch embed: Conv3d(k=s=patch) == per-patch dot product
    wpe = sd["backbone.patch_embed.proj.weight"]
    C = wpe.shape[0]
    cols = x.reshape(B, -1, D, pd, H, ph, W, pw).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, D * H * W, -1)
    t = cols @ wpe.reshape(C, -1).t() + sd["backbone.patch_embed.proj.bias"]
    t = _ln(t, sd["backbone.patch_embed.norm.weight"], sd["backbone.patch_embed.norm.bias"])
    if taps is not None:
        taps["embed"] = t.detach().clone()

    nblk = sum(depths)
    dpr = [drop_path_rate * i / max(nblk - 1, 1) for i in range(nblk)]
    shift = tuple(w // 2 for w in window)
    bi = 0
    for s, depth in enumerate(depths):
        grid = padded_grid((D, H, W), window)
        Dp, Hp, Wp = grid
        t = F.pad(t.reshape(B, D, H, W, C), (0, 0, 0, Wp - W, 0, Hp - H, 0, Dp - D)).reshape(B, Dp * Hp * Wp, C)
        for i in range(depth):
            pre = f"backbone.layers.{s}.blocks.{i}."
            y = _ln(t, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
            a = window_attention(y, sd, pre + "attn.", heads[s], grid, window, shift, shifted=(i % 2 == 1))
            t = t + _droppath(a, masks, dpr[bi], training)
            y = _ln(t, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
            m = _gelu(y @ sd[pre + "mlp.0.weight"].t() + sd[pre + "mlp.0.bias"])
            m = m @ sd[pre + "mlp.3.weight"].t() + sd[pre + "mlp.3.bias"]
            t = t + _droppath(m, masks, dpr[bi], training)
            bi += 1
        t = t.reshape(B, Dp, Hp, Wp, C)[:, :D, :H, :W]
        if s < len(depths) - 1:
            pre = f"backbone.layers.{s}.downsample."
            t = F.pad(t, (0, 0, 0, W % 2, 0, H % 2, 0, D % 2))
            parts = [t[:, a::2, b::2, c::2] for (a, b, c) in
                     ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))]
            t = torch.cat(parts, -1)
            D, H, W = t.shape[1:4]
            t = _ln(t, sd[pre + "norm.weight"], sd[pre + "norm.bias"]) @ sd[pre + "reduction.weight"].t()
            C = 2 * C
        t = t.reshape(B, D * H * W, C)
        if taps is not None:
            taps[f"stage{s}"] = t.detach().clone()
    t = _ln(t, sd["backbone.norm.weight"], sd["backbone.norm.bias"]).mean(1)
    return t @ sd["head.weight"].t() + sd["head.bias"]


def vit_forward(sd: Dict[str, Tensor], x: Tensor, *, patch=(16, 16, 16), heads=6, dim_head=64,
                depth=12, taps: Optional[dict] = None) -> Tensor:
    """ViT-3D pre-norm forward (models/vit_3d.py:438-457).  x: [B,c,D,H,W]."""
    B, c = x.shape[:2]
    p1, p2, p3 = patch
    d, h, w = x.shape[2] // p1, x.shape[3] // p2, x.shape[4] // p3
    # 'b c (d p1) (h p2) (w p3) -> b (d h w) (p1 p2 p3 c)'
    t = x.reshape(B, c, d, p1, h, p2, w, p3).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(B, d * h * w, -1)
    t = _ln(t, sd["to_patch_embedding.1.weight"], sd["to_patch_embedding.1.bias"])
    t = t @ sd["to_patch_embedding.2.weight"].t() + sd["to_patch_embedding.2.bias"]
    t = _ln(t, sd["to_patch_embedding.3.weight"], sd["to_patch_embedding.3.bias"])
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], 1)
    t = t + sd["pos_embedding"][:, : t.shape[1]]
    if taps is not None:
        taps["embed"] = t.detach().clone()
    inner = heads * dim_head
    for i in range(depth):
        pa, pf = f"transformer.layers.{i}.0.", f"transformer.layers.{i}.1.net."
        y = _ln(t, sd[pa + "norm.weight"], sd[pa + "norm.bias"])
        qkv = (y @ sd[pa + "to_qkv.weight"].t()).reshape(B, -1, 3, heads, dim_head)
        q, k, v = (qkv[:, :, j].permute(0, 2, 1, 3) for j in range(3))
        s = (q @ k.transpose(-1, -2)) * dim_head ** -0.5
        o = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, -1, inner)
        t = t + (o @ sd[pa + "to_out.0.weight"].t() + sd[pa + "to_out.0.bias"])
        y = _ln(t, sd[pf + "0.weight"], sd[pf + "0.bias"])
        m = _gelu(y @ sd[pf + "1.weight"].t() + sd[pf + "1.bias"])
        t = t + (m @ sd[pf + "4.weight"].t() + sd[pf + "4.bias"])
    t = _ln(t[:, 0], sd["mlp_head.0.weight"], sd["mlp_head.0.bias"])
    return t @ sd["mlp_head.1.weight"].t() + sd["mlp_head.1.bias"]


def soft_target_ce(logits: Tensor, target: Tensor, smoothing: float = 0.0) -> Tensor:
    """regularization/label_smoothing.py:33-77 with reduction='mean'."""
    k = logits.shape[-1]
    tgt = target.to(logits.dtype)
    if smoothing > 0.0:
        tgt = tgt * (1.0 - smoothing) + smoothing / k
    return -(tgt * torch.log_softmax(logits, -1)).sum(-1).mean()


# ----------------------------------------------------------------------------
# SAM / EMA on flat numpy arrays
# ----------------------------------------------------------------------------

def sam_grad_norm(grads: Sequence[np.ndarray], params: Optional[Sequence[np.ndarray]] = None,
                  adaptive: bool = False) -> float:
    """regularization/sam.py:122-155: 2-norm of the per-tensor 2-norms, skipping
    tensors whose norm is NaN/Inf; 1e-12 if nothing is left or the total is not finite."""
    norms = []
    for i, g in enumerate(grads):
        v = (np.abs(params[i]) * g) if adaptive else g
        n = float(np.sqrt(np.sum(v.astype(np.float64) ** 2)))
        if np.isfinite(n):
            norms.append(n)
    if not norms:
        return 1e-12
    tot = float(np.sqrt(np.sum(np.square(norms))))
    return tot if np.isfinite(tot) else 1e-12


def sam_first_step(params: List[np.ndarray], grads: Sequence[np.ndarray], rho: float,
                   adaptive: bool = False) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """regularization/sam.py:38-75.  Returns (perturbed params, saved old params)."""
    n = sam_grad_norm(grads, params, adaptive)
    old = [p.copy() for p in params]
    if not np.isfinite(n) or n == 0.0:
        return [p.copy() for p in params], old
    scale = np.float32(rho / (n + 1e-12))
    out = []
    for p, g in zip(params, grads):
        e = ((p * p) if adaptive else np.float32(1.0)) * g * scale
        out.append(p.copy() if not np.all(np.isfinite(e)) else (p + e).astype(p.dtype))
    return out, old


def ema_weights(k: int, decay: float) -> List[float]:
    """utils/ema.py:91-95: weights for the k retained snapshots, oldest first."""
    w = [decay ** i for i in range(k)][::-1]
    s = sum(w)
    return [v / s for v in w]


def ema_average(states: Sequence[np.ndarray], decay: float) -> np.ndarray:
    """utils/ema.py:97-108 for one floating tensor: zero, then add w_i*state_i
    oldest→newest in the tensor's own dtype (fp32)."""
    w = ema_weights(len(states), decay)
    acc = np.zeros_like(states[0])
    for s, wi in zip(states, w):
        acc = (acc + s * np.asarray(wi, dtype=acc.dtype)).astype(acc.dtype)
    return acc
