"""Swin-3D forward/backward built from the vsn_b200 kernels.

The residual stream is an fp32 token matrix `[B*Dp*Hp*Wp, C]` in natural (b,d,h,w) order on the
*padded* stage grid (the reference pads once per stage and keeps the padded tokens alive,
models/swin_transformer_3d.py:457-461,508).  GEMM operands are bf16, accumulation / LayerNorm /
softmax statistics are fp32.  Each autograd.Function below is one unit of the reference's module tree
with its backward written by hand on top of the same C ABI (no autograd through torch ops).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops

F32, BF16 = torch.float32, ops.BF16

# dtype of the two activation gradients that feed the LayerNorm backward kernels (outputs of the qkv / fc1 dgrad GEMMs).
# fp32 keeps the LayerNorm parameter gradients and the residual-stream gradient free of one bf16 rounding.
LN_DY_DTYPE = BF16


def _contig_f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != F32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------
# bf16 shadow of the GEMM weights (one flat buffer, refreshed by one kernel per forward)
# --------------------------------------------------------------------------------------
class WeightShadow:
    """bf16 copies of the fp32 master weights that feed tensor-core GEMMs.

    The copies live in one flat buffer and are refreshed with a single multi-tensor launch at the start
    of every model forward, so they can never be stale with respect to optimiser / SAM / EMA writes
    (which bypass torch's version counters when they come from our own kernels)."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = list(params)
        self._key = None
        self._flat = None
        self._views: List[torch.Tensor] = []
        self._table = None
        self.hold = False       # True: forward() does not refresh (the caller does, once per optimiser pass)

    def refresh(self, force: bool = False) -> None:
        from .optim import MultiTensorTable
        key = tuple((p.data_ptr(), p.numel()) for p in self.params)
        if key == self._key and self.hold and not force:
            return      # the owner (train.TrainStep) refreshes once per optimiser pass: masters unchanged since
        if key != self._key:
            dev = self.params[0].device
            sizes = [(p.numel() + 7) // 8 * 8 for p in self.params]   # keep every view 16-byte aligned
            self._flat = torch.empty(sum(sizes), device=dev, dtype=BF16)
            self._views, off = [], 0
            for p, s in zip(self.params, sizes):
                self._views.append(self._flat[off: off + p.numel()].view(p.shape[0], -1))
                off += s
            self._table = MultiTensorTable([p.detach() for p in self.params], dev)
            self._dst_ptrs = self._table.ptr_array(self._views)
            self._src_ptrs = self._table.ptr_array([p.detach() for p in self.params])
            self._key = key
        self._table.cast_bf16(self._src_ptrs, self._dst_ptrs)

    def view(self, i: int) -> torch.Tensor:
        return self._views[i]


# --------------------------------------------------------------------------------------
# parameter gradients written in place by the kernels
# --------------------------------------------------------------------------------------
class GradSink:
    """Every parameter-gradient kernel of the path accumulates (+=).  While the sink is enabled
    (`train.TrainStep` does that) the Functions below accumulate straight into the parameters' persistent `.grad`
    storage (views into the flat arena of `ddp.GradAllReduce`) and return no parameter gradients to autograd: no
    per-block zero fills, no per-parameter `grad += new` kernels, and the micro-batches of an optimiser step sum up
    in place -- the accumulation the reference gets from `loss.backward()` under `model.no_sync()`
    (train/train_transformer.py:1131-1137).  `notify`, if set, is told which gradients a unit has just completed
    (the data-parallel exchange reduces a bucket as soon as its last gradient is in).  Disabled, every backward
    returns its gradients to autograd as usual (the reference trainer, torch DDP, `w.watch` hooks)."""

    enabled = False
    notify = None       # callable(list of .grad views) or None

    @classmethod
    def wrap(cls, p):
        """Model code hands every parameter to the Functions through this.  Enabled: a fresh leaf that shares the
        parameter's storage and carries the `.grad` view the kernels accumulate into.  (A fresh leaf per forward also
        keeps autograd's per-parameter AccumulateGrad nodes -- which remember the stream they were created on and
        stay alive as long as any earlier graph does -- out of a CUDA-graph capture: a node from an eager pass would
        make the engine record an event on the default stream in the middle of the capture.)"""
        if not cls.enabled or p is None:
            return p
        g = p.grad
        if g is None or g.dtype != F32 or not g.is_contiguous() or g.shape != p.shape:
            raise RuntimeError("GradSink needs a persistent contiguous fp32 .grad on every parameter "
                               "(ddp.GradAllReduce provides them)")
        q = p.detach().requires_grad_(True)
        q._vsn_sink = g
        return q

    @classmethod
    def destinations(cls, *params):
        """Called in forward: the `.grad` views the unit's backward will accumulate into, or None."""
        if not cls.enabled:
            return None
        out = []
        for p in params:
            if p is None:
                out.append(None)
                continue
            g = getattr(p, "_vsn_sink", None)
            if g is None:
                raise RuntimeError("GradSink is enabled but a parameter reached a Function without GradSink.wrap()")
            out.append(g)
        return out

    @classmethod
    def done(cls, grads) -> None:
        if cls.notify is not None:
            cls.notify([g for g in grads if g is not None])


# --------------------------------------------------------------------------------------
# per-unit autograd Functions
# --------------------------------------------------------------------------------------
@dataclass
class BlockCfg:
    heads: int
    hd: int
    geom: Optional[ops.WindowGeom]  # window geometry incl. shift / mask flag; None = dense (ViT) attention
    tokens_per_sample: int
    S: int = 0                      # sequences (windows or samples) and tokens per sequence
    N: int = 0
    w16: Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor] = None   # qkv, proj, fc1, fc2 (bf16)
    scale1: Optional[torch.Tensor] = None   # DropPath keep/(1-p) per sample for the attention branch
    scale2: Optional[torch.Tensor] = None   # ... for the MLP branch
    # backward side channel: the block that consumes this block's input gradient (the previous block of the stage)
    # starts by casting it to bf16 scaled by ITS MLP DropPath factor; when `emit_for_prev` is set this block's last
    # LayerNorm-backward kernel writes that bf16 copy as a second output and `take_from_next` tells the consumer to
    # pick it up instead of running the cast kernel (saves one read of the fp32 gradient and a launch per block).
    emit_for_prev: bool = False
    prev_scale2: Optional[torch.Tensor] = None
    take_from_next: bool = False


_SIDE: dict = {}     # fp32 gradient data_ptr -> its bf16 copy pre-scaled for the consuming block (see BlockCfg)


class SwinBlockFn(torch.autograd.Function):
    """One pre-norm transformer block: SwinTransformerBlock.forward (models/swin_transformer_3d.py:328-380) with
    window attention, or the ViT block (models/vit_3d.py:129-142,237-254) with dense attention (geom None,
    no qkv bias, no bias table)."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkv_w, qkv_b, table, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b,
                cfg: BlockCfg):
        x = _contig_f32(x)
        wq, wp, w1, w2 = cfg.w16
        g, tps = cfg.geom, cfg.tokens_per_sample
        S, N = (g.S, g.N) if g is not None else (cfg.S, cfg.N)
        y1, mean1, rstd1 = ops.layernorm_fwd(x, n1w, n1b)
        qkv = ops.linear_fwd(y1, wq, qkv_b)
        o, lse = ops.attn_fwd(qkv, cfg.heads, cfg.hd, S=S, N=N, scale=cfg.hd ** -0.5, geom=g, table=table)
        x1 = ops.linear_fwd(o, wp, proj_b, out_dtype=F32, resid=x, row_scale=cfg.scale1, rows_per_group=tps)
        y2, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, n2b)
        h = torch.empty((x.shape[0], w1.shape[0]), device=x.device, dtype=BF16)
        a = ops.linear_fwd(y2, w1, fc1_b, gelu_aux=h)
        x2 = ops.linear_fwd(a, w2, fc2_b, out_dtype=F32, resid=x1, row_scale=cfg.scale2, rows_per_group=tps)
        ctx.cfg = cfg
        ctx.has_qkv_bias = qkv_b is not None
        ctx.has_table = table is not None
        ctx.sink = GradSink.destinations(qkv_w, proj_w, fc1_w, fc2_w, qkv_b, proj_b, fc1_b, fc2_b, n1w, n1b, n2w, n2b,
                                         table)
        ctx.save_for_backward(x, mean1, rstd1, y1, qkv, o, lse, x1, mean2, rstd2, y2, h, a, n1w, n2w,
                              table if table is not None else n1w)
        ctx.shapes = (qkv_w.shape, proj_w.shape, fc1_w.shape, fc2_w.shape)
        return x2

    @staticmethod
    def backward(ctx, g):
        cfg: BlockCfg = ctx.cfg
        x, mean1, rstd1, y1, qkv, o, lse, x1, mean2, rstd2, y2, h, a, n1w, n2w, table = ctx.saved_tensors
        wq, wp, w1, w2 = cfg.w16
        geom, tps = cfg.geom, cfg.tokens_per_sample
        S, N = (geom.S, geom.N) if geom is not None else (cfg.S, cfg.N)
        if not ctx.has_table:
            table = None
        dev = x.device
        C = x.shape[1]
        g = _contig_f32(g)
        if ctx.sink is not None:
            # in-place mode: the kernels accumulate into the parameters' own .grad storage
            bufs = [b.view(-1) if b is not None and b.dim() == 0 else b for b in ctx.sink]
            if not ctx.has_qkv_bias:
                bufs[4] = None
            release = False
        else:
            # all parameter-gradient accumulators of the block come from ONE zero-filled buffer (one fill kernel
            # instead of 13); every kernel below accumulates (+=) into its slice
            shapes = [*ctx.shapes, (ctx.shapes[0][0],), (C,), (w1.shape[0],), (C,), (C,), (C,), (C,), (C,)]
            if table is not None:
                shapes.append(tuple(table.shape))
            bufs = zeros_like_shapes(shapes, dev)
            release = True
        (d_qkv_w, d_proj_w, d_fc1_w, d_fc2_w, d_qkv_b, d_proj_b, d_fc1_b, d_fc2_b, d_n1w, d_n1b, d_n2w, d_n2b,
         *rest) = bufs
        d_table = rest[0] if table is not None else None
        # ---- MLP branch: x2 = x1 + s2 * (W2 gelu(W1 LN2(x1) + b1) + b2)
        gs = _SIDE.pop(g.data_ptr(), None) if cfg.take_from_next else None
        if gs is None or gs.shape != g.shape:
            gs = ops.cast_rows_bf16(g, cfg.scale2, tps)
        ops.linear_wgrad(gs, a, d_fc2_w, dbias=d_fc2_b)
        dh = ops.linear_dgrad(gs, w2, gelu_aux=h)
        ops.linear_wgrad(dh, y2, d_fc1_w, dbias=d_fc1_b)
        dy2 = ops.linear_dgrad(dh, w1, out_dtype=LN_DY_DTYPE)
        g1, g1s = ops.layernorm_bwd(dy2, x1, mean2, rstd2, n2w, resid_grad=g, want_bf16=True, row_scale=cfg.scale1,
                                    rows_per_group=tps, dgamma=d_n2w, dbeta=d_n2b)
        # ---- attention branch: x1 = x + s1 * (Wp attn(Wq LN1(x) + bq) + bp)
        ops.linear_wgrad(g1s, o, d_proj_w, dbias=d_proj_b)
        do = ops.linear_dgrad(g1s, wp)
        dqkv = ops.attn_bwd(qkv, o, do, lse, cfg.heads, cfg.hd, S=S, N=N, scale=cfg.hd ** -0.5, geom=geom,
                            table=table, dtable=d_table)
        ops.linear_wgrad(dqkv, y1, d_qkv_w, dbias=d_qkv_b if ctx.has_qkv_bias else None)
        dy1 = ops.linear_dgrad(dqkv, wq, out_dtype=LN_DY_DTYPE)
        g0, g0b = ops.layernorm_bwd(dy1, x, mean1, rstd1, n1w, resid_grad=g1, dx_out=g1, dgamma=d_n1w, dbeta=d_n1b,
                                    want_bf16=cfg.emit_for_prev, row_scale=cfg.prev_scale2, rows_per_group=tps)
        if cfg.emit_for_prev:
            _SIDE.clear()                       # at most one pending hand-over
            _SIDE[g0.data_ptr()] = g0b
        if not release:
            GradSink.done(ctx.sink)
            return (g0,) + (None,) * 14
        return (g0, d_n1w, d_n1b, d_qkv_w, d_qkv_b if ctx.has_qkv_bias else None, d_table, d_proj_w, d_proj_b,
                d_n2w, d_n2b, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b, None)


class PatchEmbedFn(torch.autograd.Function):
    """PatchEmbed3D.forward: pad + Conv3d(k=s=patch) + LayerNorm (models/swin_transformer_3d.py:532-543)."""

    @staticmethod
    def forward(ctx, vol, conv_w, conv_b, nw, nb, w16, patch):
        rows = ops.patch_gather(vol, patch)                                    # bf16 [T, pd*ph*pw]
        y = ops.linear_fwd(rows, w16, conv_b, out_dtype=F32)                   # conv output, fp32 [T, C]
        ctx.sink = GradSink.destinations(conv_w, conv_b, nw, nb)
        if nw is None:
            ctx.has_norm = False
            ctx.save_for_backward(rows)
            ctx.wshape = conv_w.shape
            return y
        x0, mean, rstd = ops.layernorm_fwd(y, nw, nb, out_dtype=F32)
        ctx.has_norm = True
        ctx.save_for_backward(rows, y, mean, rstd, nw)
        ctx.wshape = conv_w.shape
        return x0

    @staticmethod
    def backward(ctx, g):
        g = _contig_f32(g)
        dev = g.device
        C = g.shape[1]
        if ctx.sink is not None:
            d_w, d_b, d_nw, d_nb = ctx.sink
            d_w = d_w.view(C, -1)
        else:
            d_w, d_b, d_nw, d_nb = zeros_like_shapes([(C, ctx.wshape.numel() // C), (C,), (C,), (C,)], dev)
        if ctx.has_norm:
            rows, y, mean, rstd, nw = ctx.saved_tensors
            _, dyb = ops.layernorm_bwd(g, y, mean, rstd, nw, want_dx=False, want_bf16=True, dgamma=d_nw, dbeta=d_nb)
        else:
            (rows,) = ctx.saved_tensors
            d_nw = d_nb = None
            dyb = ops.cast_rows_bf16(g)
        ops.linear_wgrad(dyb, rows, d_w, dbias=d_b)
        if ctx.sink is not None:
            GradSink.done(ctx.sink)
            return (None,) * 7
        return None, d_w.view(ctx.wshape), d_b, d_nw, d_nb, None, None


class GridCopyFn(torch.autograd.Function):
    """Zero-pad to the window multiple / crop back (models/swin_transformer_3d.py:457-461,508)."""

    @staticmethod
    def forward(ctx, x, sdims, ddims, B):
        ctx.meta = (sdims, ddims, B, x.shape[1])
        return ops.grid_copy(_contig_f32(x), sdims, ddims, B, x.shape[1])

    @staticmethod
    def backward(ctx, g):
        sdims, ddims, B, C = ctx.meta
        return ops.grid_copy(_contig_f32(g), ddims, sdims, B, C), None, None, None


class PatchMergeFn(torch.autograd.Function):
    """Crop + PatchMerging: gather 2x2x2 neighbours, LayerNorm(8C), Linear(8C,2C,bias=False)
    (models/swin_transformer_3d.py:508,553-572).  x lives on the padded stage grid."""

    @staticmethod
    def forward(ctx, x, nw, nb, red_w, w16, pdims, rdims, B, prev_cfg=None):
        x = _contig_f32(x)
        C = x.shape[1]
        fused = C in ops.MERGE_LN_WIDTHS          # gather fused into the LayerNorm: the merged fp32 row never exists
        # backward hand-over to the last block of the stage (see BlockCfg): the fused backward also writes the bf16
        # copy of dx scaled by that block's MLP DropPath factor, which the block would otherwise cast itself
        ctx.prev_cfg = prev_cfg if fused else None
        if ctx.prev_cfg is not None:
            prev_cfg.take_from_next = True
        if fused:
            y, mean, rstd = ops.merge_ln_fwd(x, nw, nb, pdims, rdims, B, C)
            xg = x
        else:
            xg = ops.merge_gather(x, pdims, rdims, B, C)
            y, mean, rstd = ops.layernorm_fwd(xg, nw, nb)
        out = ops.linear_fwd(y, w16, None, out_dtype=F32)
        ctx.meta = (pdims, rdims, B, C, red_w.shape, w16, fused)
        ctx.sink = GradSink.destinations(nw, nb, red_w)
        ctx.save_for_backward(xg, mean, rstd, y, nw)
        return out

    @staticmethod
    def backward(ctx, g):
        pdims, rdims, B, C, wshape, w16, fused = ctx.meta
        xg, mean, rstd, y, nw = ctx.saved_tensors
        dev = g.device
        gb = ops.cast_rows_bf16(_contig_f32(g))
        if ctx.sink is not None:
            d_nw, d_nb, d_w = ctx.sink
        else:
            d_nw, d_nb, d_w = zeros_like_shapes([(8 * C,), (8 * C,), tuple(wshape)], dev)
        ops.linear_wgrad(gb, y, d_w)
        dy = ops.linear_dgrad(gb, w16)
        if fused:
            pc = ctx.prev_cfg
            dx, dxb = ops.merge_ln_bwd(dy, xg, mean, rstd, nw, d_nw, d_nb, pdims, rdims, B, C,
                                       want_bf16=pc is not None, row_scale=pc.scale2 if pc is not None else None)
            if pc is not None:
                _SIDE.clear()
                _SIDE[dx.data_ptr()] = dxb
        else:
            dxg, _ = ops.layernorm_bwd(dy, xg, mean, rstd, nw, dgamma=d_nw, dbeta=d_nb)
            dx = ops.merge_scatter(dxg, pdims, rdims, B, C)
        if ctx.sink is not None:
            GradSink.done(ctx.sink)
            return (dx,) + (None,) * 8
        return dx, d_nw, d_nb, d_w, None, None, None, None, None


class NormPoolHeadFn(torch.autograd.Function):
    """backbone.norm + AdaptiveAvgPool3d(1) + flatten + head Linear
    (models/swin_transformer_3d.py:692-698,758-761).  x: [B*T, F] on the real (cropped) grid."""

    @staticmethod
    def forward(ctx, x, nw, nb, head_w, head_b, B, T):
        x = _contig_f32(x)
        Fd = x.shape[1]
        y, mean, rstd = ops.layernorm_fwd(x, nw, nb, out_dtype=F32)
        pooled = ops.token_mean(y, B, T, Fd)
        ctx.meta = (B, T, Fd, head_w is not None, head_b is not None)
        ctx.sink = GradSink.destinations(nw, nb, head_w, head_b)
        if head_w is None:
            ctx.save_for_backward(x, mean, rstd, nw)
            return pooled
        logits = ops.head_fwd(pooled, head_w, head_b)
        ctx.save_for_backward(x, mean, rstd, nw, pooled, head_w)
        return logits

    @staticmethod
    def backward(ctx, g):
        B, T, Fd, has_head, has_bias = ctx.meta
        g = _contig_f32(g)
        dev = g.device
        sink = ctx.sink
        if has_head:
            x, mean, rstd, nw, pooled, head_w = ctx.saved_tensors
            if sink is not None:
                d_hw = sink[2]
                d_hb = sink[3] if has_bias else torch.zeros(head_w.shape[0], device=dev, dtype=F32)
            else:
                d_hw, d_hb = zeros_like_shapes([tuple(head_w.shape), (head_w.shape[0],)], dev)
            dfeat = ops.head_bwd(g, pooled, head_w, d_hw, d_hb)
        else:
            x, mean, rstd, nw = ctx.saved_tensors
            d_hw = d_hb = None
            dfeat = g
        dy = ops.token_mean_bwd(dfeat, B, T, Fd)
        if sink is not None:
            d_nw, d_nb = sink[0], sink[1]
        else:
            d_nw, d_nb = zeros_like_shapes([(Fd,), (Fd,)], dev)
        dx, _ = ops.layernorm_bwd(dy, x, mean, rstd, nw, dgamma=d_nw, dbeta=d_nb)
        if sink is not None:
            GradSink.done(sink)
            return (dx,) + (None,) * 6
        return dx, d_nw, d_nb, d_hw, (d_hb if has_bias else None), None, None


# --------------------------------------------------------------------------------------
# helpers shared by the drop-in modules
# --------------------------------------------------------------------------------------
def zeros_like_shapes(shapes, device) -> List[torch.Tensor]:
    """fp32 zero tensors of the given shapes carved out of a single zero-filled allocation (16-byte aligned slices)."""
    sizes = [1 if len(s) == 0 else int(torch.Size(s).numel()) for s in shapes]
    offs, tot = [], 0
    for n in sizes:
        offs.append(tot)
        tot += (n + 3) // 4 * 4
    flat = torch.zeros(tot, device=device, dtype=F32)
    return [flat[o: o + n].view(tuple(s)) for o, n, s in zip(offs, sizes, shapes)]


def padded_dims(real: Sequence[int], window: Sequence[int]) -> Tuple[int, int, int]:
    return tuple((r + w - 1) // w * w for r, w in zip(real, window))


def relative_position_index(window: Sequence[int]) -> torch.Tensor:
    """int64 [N,N] buffer kept only for state_dict compatibility (the kernels use the closed form
    idx = lin(i) - lin(j) + offset; models/swin_transformer_3d.py:132-152)."""
    wd, wh, ww = window
    t = torch.arange(wd * wh * ww)
    lin = (t // (wh * ww)) * ((2 * wh - 1) * (2 * ww - 1)) + ((t // ww) % wh) * (2 * ww - 1) + (t % ww)
    off = (wd - 1) * (2 * wh - 1) * (2 * ww - 1) + (wh - 1) * (2 * ww - 1) + (ww - 1)
    return (lin[:, None] - lin[None, :] + off).to(torch.int64)


def droppath_scale(p: float, B: int, device, training: bool, forced=None) -> Optional[torch.Tensor]:
    """Per-sample keep/(1-p) factors (timm DropPath semantics); None = identity."""
    if p == 0.0 or not training:
        return None
    keep = 1.0 - p
    if forced is not None:
        m = next(forced).to(device=device, dtype=F32)
    else:
        m = torch.empty(B, device=device, dtype=F32).bernoulli_(keep)
    return (m / keep).contiguous()
