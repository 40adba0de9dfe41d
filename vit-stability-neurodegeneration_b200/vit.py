"""ViT-3D specific autograd Functions (patch embedding with cls/pos assembly, cls-token head).
The transformer blocks reuse `swin.SwinBlockFn` in dense-attention mode."""
from __future__ import annotations

import torch

from . import ops
from .swin import GradSink, _contig_f32, zeros_like_shapes

F32, BF16 = torch.float32, ops.BF16


class ViTEmbedFn(torch.autograd.Function):
    """to_patch_embedding (Rearrange -> LayerNorm(P) -> Linear(P, C) -> LayerNorm(C)), cls token concat and
    `x += pos_embedding` (models/vit_3d.py:364-374,445-449)."""

    @staticmethod
    def forward(ctx, vol, ln1w, ln1b, lin_w, lin_b, ln2w, ln2b, cls, pos, w16, patch):
        B = vol.shape[0]
        fused = ops.patch_ln_supported(patch, vol)
        if fused:            # Rearrange + LayerNorm(P) in one pass over the volume: no [B*T, P] fp32 rows
            vol = vol.contiguous()
            y1, m1, r1 = ops.patch_ln_fwd(vol, patch, ln1w, ln1b)
            rows = vol
        else:
            rows = ops.patch_gather(vol, patch, out_dtype=F32)             # [B*T, P]
            y1, m1, r1 = ops.layernorm_fwd(rows, ln1w, ln1b)
        T = y1.shape[0] // B
        z = ops.linear_fwd(y1, w16, lin_b, out_dtype=F32)
        e, m2, r2 = ops.layernorm_fwd(z, ln2w, ln2b, out_dtype=F32)
        C = z.shape[1]
        if pos.shape[1] < T + 1:
            raise ValueError(f"pos_embedding has {pos.shape[1]} positions, input needs {T + 1}")
        x = ops.vit_assemble(e, cls.reshape(-1), pos.reshape(pos.shape[1], C), B, T, C)
        ctx.meta = (B, T, C, w16, lin_w.shape, cls.shape, pos.shape, tuple(patch) if fused else None)
        ctx.sink = GradSink.destinations(ln1w, ln1b, lin_w, lin_b, ln2w, ln2b, cls, pos)
        ctx.save_for_backward(rows, m1, r1, y1, z, m2, r2, ln1w, ln2w)
        return x

    @staticmethod
    def backward(ctx, g):
        B, T, C, w16, wshape, cls_shape, pos_shape, fused_patch = ctx.meta
        rows, m1, r1, y1, z, m2, r2, ln1w, ln2w = ctx.saved_tensors
        g = _contig_f32(g)
        dev = g.device
        P = y1.shape[1]
        if ctx.sink is not None:
            d_ln1w, d_ln1b, d_lin_w, d_lin_b, d_ln2w, d_ln2b, d_cls, d_pos = ctx.sink
            d_cls, d_pos = d_cls.view(C), d_pos.view(pos_shape[1], C)
        else:
            d_ln1w, d_ln1b, d_lin_w, d_lin_b, d_ln2w, d_ln2b, d_cls, d_pos = zeros_like_shapes(
                [(P,), (P,), tuple(wshape), (C,), (C,), (C,), (C,), (pos_shape[1], C)], dev)
        ops.vit_assemble_bwd(g, d_cls, d_pos, B, T, C)
        # gradient wrt the patch embeddings = rows 1.. of every sample (a shifted crop of the token axis)
        demb = ops.grid_copy(g.view(-1)[C:], (T + 1, 1, 1), (T, 1, 1), B, C)
        ops.ln_param_grad(demb, z, m2, r2, d_ln2w, d_ln2b)
        _, dzb = ops.layernorm_bwd(demb, z, m2, r2, ln2w, want_dx=False, want_bf16=True)
        ops.linear_wgrad(dzb, y1, d_lin_w, dbias=d_lin_b)
        dy1 = ops.linear_dgrad(dzb, w16)
        if fused_patch is not None:
            ops.patch_ln_param_grad(dy1, rows, fused_patch, m1, r1, d_ln1w, d_ln1b)
        else:
            ops.ln_param_grad(dy1, rows, m1, r1, d_ln1w, d_ln1b)
        if ctx.sink is not None:
            GradSink.done(ctx.sink)
            return (None,) * 11
        return (None, d_ln1w, d_ln1b, d_lin_w, d_lin_b, d_ln2w, d_ln2b, d_cls.view(cls_shape), d_pos.view(pos_shape),
                None, None)


class ViTHeadFn(torch.autograd.Function):
    """cls (or mean) pooling + mlp_head = LayerNorm -> Linear (models/vit_3d.py:454-457)."""

    @staticmethod
    def forward(ctx, x, nw, nb, head_w, head_b, B, N, pool):
        x = _contig_f32(x)
        C = x.shape[1]
        if pool == "cls":
            feat_in = x.view(B, N, C)[:, 0]                      # strided rows, no copy
        else:
            feat_in = ops.token_mean(x, B, N, C)
        y, mean, rstd = ops.layernorm_fwd(feat_in, nw, nb, out_dtype=F32)
        logits = ops.head_fwd(y, head_w, head_b)
        ctx.meta = (B, N, C, pool)
        ctx.sink = GradSink.destinations(nw, nb, head_w, head_b)
        ctx.save_for_backward(x if pool == "cls" else feat_in, mean, rstd, nw, y, head_w)
        return logits

    @staticmethod
    def backward(ctx, g):
        B, N, C, pool = ctx.meta
        xin, mean, rstd, nw, y, head_w = ctx.saved_tensors
        g = _contig_f32(g)
        dev = g.device
        if ctx.sink is not None:
            d_nw, d_nb, d_hw, d_hb = ctx.sink
        else:
            d_nw, d_nb, d_hw, d_hb = zeros_like_shapes([(C,), (C,), tuple(head_w.shape), (head_w.shape[0],)], dev)
        dy = ops.head_bwd(g, y, head_w, d_hw, d_hb)
        if pool == "cls":
            feat_in = xin.view(B, N, C)[:, 0]
            ops.ln_param_grad(dy, feat_in, mean, rstd, d_nw, d_nb)
            dx = torch.zeros((B * N, C), device=dev, dtype=F32)
            ops.layernorm_bwd(dy, feat_in, mean, rstd, nw, dx_out=dx.view(B, N, C)[:, 0])
        else:
            ops.ln_param_grad(dy, xin, mean, rstd, d_nw, d_nb)
            dfeat, _ = ops.layernorm_bwd(dy, xin, mean, rstd, nw)
            dx = ops.token_mean_bwd(dfeat, B, N, C)
        if ctx.sink is not None:
            GradSink.done(ctx.sink)
            return (dx,) + (None,) * 7
        return dx, d_nw, d_nb, d_hw, d_hb, None, None, None
