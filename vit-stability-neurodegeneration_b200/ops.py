"""Tensor-level wrappers over the C ABI: output allocation and argument marshalling only.

Every function enqueues CUDA kernels from libvsn_b200.so on torch's current stream.  Nothing here
computes with torch ops; torch is the allocator and the stream owner.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib

BF16 = _lib.act_dtype()   # the 16-bit operand dtype of this process (bfloat16, or float16 under VSN_B200_PRECISION=f16)
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vsn_b200 kernels need CUDA tensors (there is no CPU fallback)")
    _lib.check_device()


# ------------------------------------------------------------------ GEMM family
def gemm(a, lda, a_mn, b, ldb, b_mn, M, N, K, out, ldo, out_kind, *, bias=None, act=0, aux=None, ldaux=0,
         resid=None, ldr=0, row_scale=None, rows_per_group=1, alpha=1.0, split_k=1, rowsum=None):
    _require_cuda(a, b, out)
    if _lib.PROFILE is not None:
        _lib.TAG = f"M{M} N{N} K{K} majors={int(a_mn)}{int(b_mn)} out={out_kind} act={act}"
        _lib.WORK = (2 * M * N * K, 2 * (M * K + N * K) + M * N * (2 if out_kind == 0 else 4)
                     + (M * N * 4 if resid is not None else 0) + (M * N * 2 if act else 0))
    _lib.call("vsn_gemm_bf16", _p(a), lda, int(a_mn), _p(b), ldb, int(b_mn), M, N, K, _p(out), ldo, out_kind,
              _p(bias), act, _p(aux), ldaux, _p(resid), ldr, _p(row_scale), rows_per_group, float(alpha), split_k,
              _p(rowsum), _stream())


def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, out_dtype=BF16,
               gelu_aux: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
               row_scale: Optional[torch.Tensor] = None, rows_per_group: int = 1,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x @ w.T (+bias) with the fused epilogues.  x [M,K] bf16, w [N,K] bf16."""
    M, K = x.shape
    N = w.shape[0]
    assert x.dtype == BF16 and w.dtype == BF16 and x.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=out_dtype)
    gemm(x, x.stride(0), 0, w, w.stride(0), 0, M, N, K, out, out.stride(0), 0 if out.dtype == BF16 else 1,
         bias=bias, act=1 if gelu_aux is not None else 0, aux=gelu_aux,
         ldaux=gelu_aux.stride(0) if gelu_aux is not None else 0,
         resid=resid, ldr=resid.stride(0) if resid is not None else 0, row_scale=row_scale,
         rows_per_group=rows_per_group)
    return out


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, *, gelu_aux: Optional[torch.Tensor] = None,
                 out_dtype=BF16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx = dy @ w  (w [N,K] read as stored: MN-major B operand).  Optional fused GELU'(aux) factor."""
    M, N = dy.shape
    K = w.shape[1]
    assert dy.dtype == BF16 and w.dtype == BF16
    if out is None:
        out = torch.empty((M, K), device=dy.device, dtype=out_dtype)
    gemm(dy, dy.stride(0), 0, w, w.stride(0), 1, M, K, N, out, out.stride(0), 0 if out.dtype == BF16 else 1,
         act=2 if gelu_aux is not None else 0, aux=gelu_aux, ldaux=gelu_aux.stride(0) if gelu_aux is not None else 0)
    return out


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, dbias: Optional[torch.Tensor] = None) -> None:
    """dw[N,K] += dy[M,N].T @ x[M,K]  (both operands MN-major, tokens split over CTAs, fp32 atomics).
    With dbias (fp32 [N]) the bias gradient dbias[n] += sum_m dy[m,n] is produced by the same GEMM."""
    M, N = dy.shape
    K = x.shape[1]
    assert dy.dtype == BF16 and x.dtype == BF16 and dw.dtype == F32 and dw.is_contiguous()
    gemm(dy, dy.stride(0), 1, x, x.stride(0), 1, N, K, M, dw, K, 2, split_k=0, rowsum=dbias)   # 0: the library picks


# ------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x: torch.Tensor, gamma, beta, *, out_dtype=BF16, eps: float = 1e-5, want_stats: bool = True):
    """x fp32 [rows, C] (row stride free).  Returns (y, mean, rstd)."""
    rows, C = x.shape
    assert x.dtype == F32 and x.stride(1) == 1
    _require_cuda(x)
    y = torch.empty((rows, C), device=x.device, dtype=out_dtype)
    mean = torch.empty((rows,), device=x.device, dtype=F32) if want_stats else None
    rstd = torch.empty((rows,), device=x.device, dtype=F32) if want_stats else None
    if _lib.PROFILE is not None:
        _lib.TAG = f"rows{rows} C{C} out={'bf16' if out_dtype == BF16 else 'f32'}"
        _lib.WORK = (0, rows * C * (4 + (2 if out_dtype == BF16 else 4)))
    _lib.call("vsn_layernorm_fwd", _p(x), x.stride(0), _p(gamma), _p(beta), _p(y), y.stride(0),
              1 if out_dtype == BF16 else 0, _p(mean), _p(rstd), rows, C, eps, _stream())
    return y, mean, rstd


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, mean, rstd, gamma, *, resid_grad=None, want_dx=True,
                  dx_out: Optional[torch.Tensor] = None, want_bf16=False, row_scale=None, rows_per_group=1,
                  dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None):
    """Returns (dx fp32 or None, dx_bf16 or None).  dx = LN'(dy) + resid_grad.  With dgamma/dbeta (fp32 [C]) the
    parameter gradients are accumulated in the same pass."""
    rows, C = x.shape
    _require_cuda(dy, x)
    dx = dx_out if dx_out is not None else (torch.empty((rows, C), device=x.device, dtype=F32) if want_dx else None)
    dxb = torch.empty((rows, C), device=x.device, dtype=BF16) if want_bf16 else None
    if _lib.PROFILE is not None:
        _lib.TAG = f"rows{rows} C{C} dy={'bf16' if dy.dtype == BF16 else 'f32'} dx={int(dx is not None)} dxb={int(dxb is not None)}"
        _lib.WORK = (0, rows * C * ((2 if dy.dtype == BF16 else 4) + 4 + (4 if resid_grad is not None else 0)
                                    + (4 if dx is not None else 0) + (2 if dxb is not None else 0)))
    _lib.call("vsn_layernorm_bwd", _p(dy), dy.stride(0), 1 if dy.dtype == BF16 else 0, _p(x), x.stride(0),
              _p(mean), _p(rstd), _p(gamma), _p(resid_grad), resid_grad.stride(0) if resid_grad is not None else 0,
              _p(dx), dx.stride(0) if dx is not None else 0, _p(dxb), C, _p(row_scale), rows_per_group,
              _p(dgamma), _p(dbeta), rows, C, _stream())
    return dx, dxb


def ln_param_grad(dy, x, mean, rstd, dgamma, dbeta) -> None:
    rows, C = x.shape
    if _lib.PROFILE is not None:
        _lib.TAG = f"rows{rows} C{C} ln dy={'bf16' if dy.dtype == BF16 else 'f32'}"
        _lib.WORK = (0, rows * C * ((2 if dy.dtype == BF16 else 4) + 4))
    _lib.call("vsn_colreduce", _p(dy), dy.stride(0), 1 if dy.dtype == BF16 else 0, _p(x), x.stride(0), _p(mean),
              _p(rstd), _p(dgamma), _p(dbeta), rows, C, _stream())


def colsum(dy: torch.Tensor, out: torch.Tensor) -> None:
    """out[c] += sum_r dy[r, c]."""
    rows, C = dy.shape
    if _lib.PROFILE is not None:
        _lib.TAG = f"rows{rows} C{C} colsum"
        _lib.WORK = (0, rows * C * (2 if dy.dtype == BF16 else 4))
    _lib.call("vsn_colreduce", _p(dy), dy.stride(0), 1 if dy.dtype == BF16 else 0, None, 0, None, None, None,
              _p(out), rows, C, _stream())


# ------------------------------------------------------------------ attention
def _geom_tensor(geom: Sequence[int], device) -> torch.Tensor:
    return torch.tensor(list(geom), dtype=torch.int32)  # host array; the ABI copies it by value


class WindowGeom:
    """Stage geometry handed to the attention kernels (host-side ints, kept alive as a ctypes array)."""

    def __init__(self, B, grid, window, shift, use_mask):
        import ctypes
        self.B, self.grid, self.window, self.shift, self.use_mask = B, tuple(grid), tuple(window), tuple(shift), use_mask
        vals = [B, *grid, *window, *shift, 1 if use_mask else 0]
        self.arr = (ctypes.c_int * 11)(*vals)
        self.S = B * (grid[0] // window[0]) * (grid[1] // window[1]) * (grid[2] // window[2])
        self.N = window[0] * window[1] * window[2]


def attn_fwd(qkv: torch.Tensor, heads: int, hd: int, *, S: int, N: int, scale: float,
             geom: Optional[WindowGeom] = None, table: Optional[torch.Tensor] = None, want_lse: bool = True):
    import ctypes
    T = qkv.shape[0]
    C = heads * hd
    assert qkv.dtype == BF16 and qkv.is_contiguous() and qkv.shape[1] == 3 * C
    _require_cuda(qkv)
    out = torch.empty((T, C), device=qkv.device, dtype=BF16)
    npad = (N + 63) // 64 * 64
    lse = torch.empty((S, heads, npad), device=qkv.device, dtype=F32) if want_lse else None
    if _lib.PROFILE is not None:
        _lib.TAG = f"S{S} N{N} heads{heads} hd{hd} win={int(geom is not None)} mask={int(bool(geom and geom.use_mask))}"
        _lib.WORK = (4 * S * N * N * C, S * N * C * 2 * 4)
    _lib.call("vsn_attn_fwd", _p(qkv), _p(out), _p(lse), S, N, heads, hd, 1 if geom is not None else 0,
              ctypes.cast(geom.arr, ctypes.c_void_p) if geom is not None else None, _p(table),
              table.shape[0] if table is not None else 0, float(scale), _stream())
    return out, lse


def attn_bwd(qkv, out, dout, lse, heads: int, hd: int, *, S: int, N: int, scale: float,
             geom: Optional[WindowGeom] = None, table: Optional[torch.Tensor] = None,
             dtable: Optional[torch.Tensor] = None) -> torch.Tensor:
    import ctypes
    T = qkv.shape[0]
    npad = (N + 63) // 64 * 64
    assert dout.dtype == BF16 and dout.is_contiguous()
    delta = torch.empty((S, heads, npad), device=qkv.device, dtype=F32)
    tc_path = geom is not None and hd == 32 and tuple(geom.window) == (6, 7, 6)   # folds the table gradient in-kernel
    if geom is None and hd == 64:
        # dense tcgen05 path (ViT-3D): the scratch argument is the fp32 dQ accumulator [T, C] (zeroed by the library)
        dense = torch.empty((T, heads * hd), device=qkv.device, dtype=F32)
    else:
        dense = torch.zeros((heads, npad, npad), device=qkv.device, dtype=F32) if (table is not None and not tc_path) else None
    # real tokens are all written by the kernels; padded-grid tokens always belong to a window too
    dqkv = torch.empty_like(qkv)
    if _lib.PROFILE is not None:
        _lib.TAG = f"S{S} N{N} heads{heads} hd{hd} win={int(geom is not None)} mask={int(bool(geom and geom.use_mask))}"
        _lib.WORK = (8 * S * N * N * heads * hd, S * N * heads * hd * 2 * 8)
    _lib.call("vsn_attn_bwd", _p(qkv), _p(out), _p(dout), _p(lse), _p(delta), _p(dqkv), _p(dense), _p(dtable), S, N,
              heads, hd, 1 if geom is not None else 0,
              ctypes.cast(geom.arr, ctypes.c_void_p) if geom is not None else None, _p(table),
              table.shape[0] if table is not None else 0, float(scale), _stream())
    return dqkv


# ------------------------------------------------------------------ layout kernels
_IN_DTYPE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}   # volume dtypes (input side), not BF16


def patch_gather(vol: torch.Tensor, patch: Sequence[int], *, out_dtype=BF16) -> torch.Tensor:
    """vol [B,1,D,H,W] (fp32/fp16/bf16, contiguous) -> rows [B*gd*gh*gw, pd*ph*pw]."""
    B, c, D, H, W = vol.shape
    if c != 1:
        raise NotImplementedError("vsn_b200 patch embedding supports in_channels=1 (the reference's MRI volumes)")
    _require_cuda(vol)
    pd, ph, pw = patch
    gd, gh, gw = -(-D // pd), -(-H // ph), -(-W // pw)
    out = torch.empty((B * gd * gh * gw, pd * ph * pw), device=vol.device, dtype=out_dtype)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} vol{D}x{H}x{W} patch{pd}x{ph}x{pw} in={vol.dtype} out={'bf16' if out_dtype == BF16 else 'f32'}"
        _lib.WORK = (0, vol.numel() * vol.element_size() + out.numel() * out.element_size())
    _lib.call("vsn_patch_gather", _p(vol), _IN_DTYPE[vol.dtype], _p(out), 1 if out_dtype == BF16 else 0, B, D, H, W,
              pd, ph, pw, _stream())
    return out


def patch_ln_supported(patch: Sequence[int], vol: torch.Tensor) -> bool:
    """Shapes the fused Rearrange + LayerNorm kernels take: a thread per patch row, four-voxel groups, and the slab of
    one row of patches (pd*ph volume rows, padded W) in shared memory."""
    P = patch[0] * patch[1] * patch[2]
    slab = patch[0] * patch[1] * (-(-vol.shape[-1] // patch[2]) * patch[2]) * vol.element_size()
    return (vol.dtype in _IN_DTYPE and patch[0] * patch[1] <= 1024 and patch[2] <= 16 and patch[2] % 4 == 0
            and P % 128 == 0 and P <= 4096 and slab <= 200 * 1024)


def patch_ln_fwd(vol: torch.Tensor, patch: Sequence[int], gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5):
    """ViT Rearrange + LayerNorm(P) in one pass (models/vit_3d.py:364-371): (y bf16 [B*T, P], mean, rstd)."""
    B, c, D, H, W = vol.shape
    if c != 1:
        raise NotImplementedError("vsn_b200 patch embedding supports in_channels=1 (the reference's MRI volumes)")
    _require_cuda(vol, gamma, beta)
    assert vol.is_contiguous()
    pd, ph, pw = patch
    T = -(-D // pd) * -(-H // ph) * -(-W // pw)
    y = torch.empty((B * T, pd * ph * pw), device=vol.device, dtype=BF16)
    mean = torch.empty(B * T, device=vol.device, dtype=F32)
    rstd = torch.empty(B * T, device=vol.device, dtype=F32)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} vol{D}x{H}x{W} patch{pd}x{ph}x{pw} in={vol.dtype} fwd"
        _lib.WORK = (0, vol.numel() * vol.element_size() + 2 * y.numel())
    _lib.call("vsn_patch_ln_fwd", _p(vol), _IN_DTYPE[vol.dtype], B, D, H, W, pd, ph, pw, _p(gamma), _p(beta), _p(y),
              _p(mean), _p(rstd), eps, _stream())
    return y, mean, rstd


def patch_ln_param_grad(dy: torch.Tensor, vol: torch.Tensor, patch: Sequence[int], mean, rstd, dgamma, dbeta) -> None:
    """dgamma / dbeta [P] += of the LayerNorm in patch_ln_fwd; xhat is re-gathered from the volume."""
    B, _, D, H, W = vol.shape
    assert dy.dtype == BF16 and dy.is_contiguous() and vol.is_contiguous()
    _require_cuda(dy, vol, mean, rstd, dgamma, dbeta)
    pd, ph, pw = patch
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} vol{D}x{H}x{W} patch{pd}x{ph}x{pw} in={vol.dtype} pgrad"
        _lib.WORK = (0, vol.numel() * vol.element_size() + 2 * dy.numel())
    _lib.call("vsn_patch_ln_param_grad", _p(dy), _p(vol), _IN_DTYPE[vol.dtype], B, D, H, W, pd, ph, pw, _p(mean),
              _p(rstd), _p(dgamma), _p(dbeta), _stream())


def grid_copy(src: torch.Tensor, sdims, ddims, B: int, C: int) -> torch.Tensor:
    dst = torch.empty((B * ddims[0] * ddims[1] * ddims[2], C), device=src.device, dtype=F32)
    if _lib.PROFILE is not None:
        overlap = B * C * min(sdims[0], ddims[0]) * min(sdims[1], ddims[1]) * min(sdims[2], ddims[2])
        _lib.TAG = f"B{B} C{C} {tuple(sdims)}->{tuple(ddims)}"
        _lib.WORK = (0, 4 * (overlap + dst.numel()))
    _lib.call("vsn_grid_copy", _p(src), *sdims, _p(dst), *ddims, B, C, _stream())
    return dst


def merge_gather(x: torch.Tensor, pdims, rdims, B: int, C: int) -> torch.Tensor:
    od = [(r + 1) // 2 for r in rdims]
    out = torch.empty((B * od[0] * od[1] * od[2], 8 * C), device=x.device, dtype=F32)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} C{C} real{tuple(rdims)} gather"
        _lib.WORK = (0, 4 * (B * C * rdims[0] * rdims[1] * rdims[2] + out.numel()))
    _lib.call("vsn_merge_gather", _p(x), *pdims, *rdims, _p(out), B, C, 0, _stream())
    return out


def merge_scatter(dout: torch.Tensor, pdims, rdims, B: int, C: int) -> torch.Tensor:
    dx = torch.zeros((B * pdims[0] * pdims[1] * pdims[2], C), device=dout.device, dtype=F32)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} C{C} real{tuple(rdims)} scatter"
        _lib.WORK = (0, 4 * (B * C * rdims[0] * rdims[1] * rdims[2] + dout.numel()))
    _lib.call("vsn_merge_gather", _p(dx), *pdims, *rdims, _p(dout), B, C, 1, _stream())
    return dx


MERGE_LN_WIDTHS = (96, 192)     # widths whose merged row (8C) a warp holds in registers: gather fused into LayerNorm


def merge_ln_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, pdims, rdims, B: int, C: int,
                 eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """PatchMerging's gather + LayerNorm(8C) in one pass (models/swin_transformer_3d.py:553-572): returns the bf16
    operand of the reduction GEMM [rows, 8C] and mean / rstd [rows]."""
    assert C in MERGE_LN_WIDTHS and x.dtype == F32 and x.is_contiguous()
    _require_cuda(x, gamma, beta)
    od = [(r + 1) // 2 for r in rdims]
    rows = B * od[0] * od[1] * od[2]
    y = torch.empty((rows, 8 * C), device=x.device, dtype=BF16)
    mean = torch.empty(rows, device=x.device, dtype=F32)
    rstd = torch.empty(rows, device=x.device, dtype=F32)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} C{C} real{tuple(rdims)} fwd"
        _lib.WORK = (0, 4 * B * C * rdims[0] * rdims[1] * rdims[2] + 2 * y.numel())
    _lib.call("vsn_merge_ln_fwd", _p(x), *pdims, *rdims, B, C, _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), eps,
              _stream())
    return y, mean, rstd


def merge_ln_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor,
                 dgamma: torch.Tensor, dbeta: torch.Tensor, pdims, rdims, B: int, C: int, *, want_bf16: bool = False,
                 row_scale: Optional[torch.Tensor] = None):
    """Backward of merge_ln_fwd: dx on the padded stage grid [B*pD*pH*pW, C]; dgamma / dbeta accumulate.  Returns
    (dx, dx_bf16 or None); dx_bf16 = dx * row_scale[sample] for the block backward that consumes dx next."""
    assert C in MERGE_LN_WIDTHS and dy.dtype == BF16 and dy.is_contiguous() and x.dtype == F32 and x.is_contiguous()
    _require_cuda(dy, x, mean, rstd, gamma, dgamma, dbeta)
    n = B * pdims[0] * pdims[1] * pdims[2]
    # every token of the real grid belongs to exactly one merged row: only a larger padded grid needs the zero fill
    alloc = torch.empty if tuple(pdims) == tuple(rdims) else torch.zeros
    dx = alloc((n, C), device=dy.device, dtype=F32)
    dxb = alloc((n, C), device=dy.device, dtype=BF16) if want_bf16 else None
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} C{C} real{tuple(rdims)} bwd dxb={int(want_bf16)}"
        _lib.WORK = (0, 2 * dy.numel() + (8 + (2 if want_bf16 else 0)) * B * C * rdims[0] * rdims[1] * rdims[2])
    _lib.call("vsn_merge_ln_bwd", _p(dy), _p(x), *pdims, *rdims, B, C, _p(mean), _p(rstd), _p(gamma), _p(dx),
              _p(dgamma), _p(dbeta), _p(dxb), _p(row_scale) if want_bf16 else None, _stream())
    return dx, dxb


def mixup(x: torch.Tensor, lam: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """out[b] = lam[b] * x[b] + (1 - lam[b]) * x[perm[b]] on fp16 volumes [B,1,D,H,W] (dataset/dataset.py:276-281)."""
    assert x.dtype == torch.float16 and x.is_contiguous() and lam.dtype == F32 and perm.dtype == torch.int32
    _require_cuda(x, lam, perm)
    B = x.shape[0]
    per = x.numel() // B
    out = torch.empty_like(x)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} elems{per}"
        _lib.WORK = (0, x.numel() * 6)
    _lib.call("vsn_mixup_f16", _p(x), _p(out), _p(lam), _p(perm), B, per, _stream())
    return out


_STATS_SCRATCH = {}


def volume_stats(x: torch.Tensor, lam: Optional[torch.Tensor] = None, perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, 2] = (mean, 1/std) of every (optionally mixed) fp16 volume: the statistics of monai's NormalizeIntensity()
    (train/train_transformer.py:1729-1752), population std, 1/std := 1 when std == 0."""
    assert x.dtype == torch.float16 and x.is_contiguous()
    _require_cuda(x)
    B = x.shape[0]
    per = x.numel() // B
    key = (x.device, B)
    if key not in _STATS_SCRATCH:
        _STATS_SCRATCH[key] = torch.zeros(2 * B, device=x.device, dtype=torch.float64)
    stats = torch.empty((B, 2), device=x.device, dtype=F32)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} elems{per} mix={int(lam is not None)}"
        _lib.WORK = (0, x.numel() * 2 * (2 if lam is not None else 1))
    _lib.call("vsn_volume_stats_f16", _p(x), _p(lam), _p(perm), B, per, _p(_STATS_SCRATCH[key]), _p(stats), _stream())
    return stats


def mixup_zscore(x: torch.Tensor, lam: Optional[torch.Tensor] = None, perm: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """MixUp (dataset/dataset.py:230-286) followed by NormalizeIntensity() of fp16 volumes [B,1,D,H,W] on the device,
    two passes over the batch (statistics, then mix + normalise + write); returns (normalised fp16 volumes, stats)."""
    if lam is not None:
        assert lam.dtype == F32 and perm is not None and perm.dtype == torch.int32
    stats = volume_stats(x, lam, perm)
    B = x.shape[0]
    per = x.numel() // B
    if out is None:
        out = torch.empty_like(x)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} elems{per} mix={int(lam is not None)}"
        _lib.WORK = (0, x.numel() * 2 * (3 if lam is not None else 2))
    _lib.call("vsn_mixup_zscore_f16", _p(x), _p(out), _p(lam), _p(perm), _p(stats), B, per, _stream())
    return out, stats


def tta_views(x: torch.Tensor, mats: torch.Tensor) -> torch.Tensor:
    """[B,1,D,H,W] fp16 volumes, [V,18] fp32 views (3x4 matrix + clamp box, see tta.py) -> [B*V,1,D,H,W] fp16
    (sample-major): every TTA view of every volume in one launch (eval/test_time_augmentation.py:221-354)."""
    assert x.dtype == torch.float16 and x.is_contiguous() and x.ndim == 5 and x.shape[1] == 1
    assert mats.dtype == F32 and mats.is_contiguous() and mats.ndim == 2 and mats.shape[1] == 18
    _require_cuda(x, mats)
    B, _, D, H, W = x.shape
    V = mats.shape[0]
    out = torch.empty((B * V, 1, D, H, W), device=x.device, dtype=torch.float16)
    if _lib.PROFILE is not None:
        _lib.TAG = f"B{B} V{V} vol{D}x{H}x{W}"
        _lib.WORK = (0, 2 * (x.numel() + out.numel()))
    _lib.call("vsn_tta_views_f16", _p(x), _p(out), _p(mats), B, V, D, H, W, _stream())
    return out


def cast_rows_bf16(src: torch.Tensor, row_scale=None, rows_per_group=1) -> torch.Tensor:
    rows, C = src.shape
    assert src.dtype == F32 and src.is_contiguous()
    dst = torch.empty((rows, C), device=src.device, dtype=BF16)
    if _lib.PROFILE is not None:
        _lib.TAG = f"rows{rows} C{C}"
        _lib.WORK = (0, rows * C * 6)
    _lib.call("vsn_cast_rows_bf16", _p(src), _p(dst), _p(row_scale), rows_per_group, rows, C, _stream())
    return dst


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert src.dtype == F32 and src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=BF16)
    _require_cuda(src)
    _lib.call("vsn_cast_bf16", _p(src), _p(dst), src.numel(), _stream())
    return dst


def token_mean(x: torch.Tensor, B: int, T: int, C: int) -> torch.Tensor:
    out = torch.empty((B, C), device=x.device, dtype=F32)
    _lib.call("vsn_token_mean", _p(x), _p(out), B, T, C, 0, _stream())
    return out


def token_mean_bwd(dout: torch.Tensor, B: int, T: int, C: int) -> torch.Tensor:
    dx = torch.empty((B * T, C), device=dout.device, dtype=F32)
    _lib.call("vsn_token_mean", _p(dout), _p(dx), B, T, C, 1, _stream())
    return dx


def head_fwd(feat: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    B, Fdim = feat.shape
    K = W.shape[0]
    _require_cuda(feat)
    out = torch.empty((B, K), device=feat.device, dtype=F32)
    _lib.call("vsn_head_fwd", _p(feat), _p(W), _p(bias), _p(out), B, K, Fdim, _stream())
    return out


def head_bwd(dlogits, feat, W, dW, db) -> torch.Tensor:
    B, Fdim = feat.shape
    K = W.shape[0]
    dfeat = torch.empty_like(feat)
    _lib.call("vsn_head_bwd", _p(dlogits), _p(feat), _p(W), _p(dfeat), _p(dW), _p(db), B, K, Fdim, _stream())
    return dfeat


def vit_assemble(emb, cls, pos, B: int, T: int, C: int) -> torch.Tensor:
    x = torch.empty((B * (T + 1), C), device=emb.device, dtype=F32)
    _lib.call("vsn_vit_assemble", _p(emb), _p(cls), _p(pos), _p(x), B, T, C, _stream())
    return x


def vit_assemble_bwd(dx, dcls, dpos, B: int, T: int, C: int) -> None:
    _lib.call("vsn_vit_assemble_bwd", _p(dx), _p(dcls), _p(dpos), B, T, C, _stream())
