"""Kernel-level parity (GPU): every C-ABI entry point against a plain torch fp32 restatement of the same
op (floating point) or the numpy oracle (index kernels, bit-exact)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import swin3d_oracle as O  # noqa: E402
from tests.helpers import rel_err  # noqa: E402

BF16_TOL = 1e-2   # bf16 operands + bf16 output rounding (2^-9 per element), fp32 accumulation


def _ops():
    import vsn_b200  # noqa: F401
    from vsn_b200 import ops
    return ops


def _bf(t):
    return t.to(torch.bfloat16)


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


# ----------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(256, 96, 64), (1000, 288, 96), (4096, 384, 96), (777, 96, 384),
                                   (252, 768, 3072), (130, 1152, 384), (8, 64, 4096), (300, 2304, 768)])
def test_gemm_forward_bias(M, N, K):
    ops = _ops()
    x, w, b = _bf(_rand(M, K, seed=1)), _bf(_rand(N, K, seed=2, scale=K ** -0.5)), _rand(N, seed=3)
    y = ops.linear_fwd(x, w, b)
    ref = x.float() @ w.float().t() + b
    assert rel_err(y.float(), ref) < BF16_TOL
    y32 = ops.linear_fwd(x, w, b, out_dtype=torch.float32)
    assert rel_err(y32, ref) < 1e-5


def test_gemm_epilogues():
    ops = _ops()
    M, N, K, rpg = 600, 384, 96, 200
    x, w, b = _bf(_rand(M, K, seed=1)), _bf(_rand(N, K, seed=2, scale=K ** -0.5)), _rand(N, seed=3)
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    y = ops.linear_fwd(x, w, b, gelu_aux=aux)
    pre = x.float() @ w.float().t() + b
    assert rel_err(aux.float(), pre) < BF16_TOL
    assert rel_err(y.float(), torch.nn.functional.gelu(aux.float())) < BF16_TOL
    # residual + per-sample scale, fp32 out
    resid, scale = _rand(M, N, seed=4), torch.tensor([1.0, 0.0, 1.25], device="cuda")
    y2 = ops.linear_fwd(x, w, b, out_dtype=torch.float32, resid=resid, row_scale=scale, rows_per_group=rpg)
    ref = resid + scale.repeat_interleave(rpg)[:, None] * pre
    assert rel_err(y2, ref) < 1e-5
    # dgrad with fused GELU'
    dy = _bf(_rand(M, K, seed=5))          # pretend gradient wrt an [M,K] output of a K<-N layer
    w2 = _bf(_rand(K, N, seed=6, scale=N ** -0.5))   # layer weight [out=K, in=N]
    dx = ops.linear_dgrad(dy, w2, gelu_aux=aux)
    a = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(a).backward(dy.float() @ w2.float())
    assert rel_err(dx.float(), a.grad) < BF16_TOL


def test_gemm_gelu_epilogue_accuracy_wide_range():
    """The packed (half2) GELU / GELU' of the epilogue against torch's exact erf GELU on pre-activations that span
    [-12, 12] (identity weights make the pre-activation equal to the input): element-wise error bounded by the bf16
    output rounding, not by the fp16 polynomial."""
    ops = _ops()
    M, K = 1024, 128
    x = torch.linspace(-12, 12, M * K, device="cuda").reshape(M, K)
    x = x[:, torch.randperm(K, device="cuda")].contiguous()
    xb = _bf(x)
    eye = _bf(torch.eye(K, device="cuda"))
    aux = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    y = ops.linear_fwd(xb, eye, None, gelu_aux=aux)
    assert torch.equal(aux, xb)
    ref = torch.nn.functional.gelu(xb.float())
    assert float((y.float() - ref).abs().max()) < 2 ** -7 * 12                     # bf16 half-ulp at |y| <= 12 plus margin
    assert float(((y.float() - ref).abs() / (ref.abs() + 1e-2)).max()) < 2e-2
    dy = _bf(torch.ones(M, K, device="cuda"))
    dx = ops.linear_dgrad(dy, eye, gelu_aux=aux)
    a = xb.float().requires_grad_(True)
    torch.nn.functional.gelu(a).sum().backward()
    assert float((dx.float() - a.grad).abs().max()) < 1.5e-2
    assert rel_err(dx.float(), a.grad) < 5e-3


@pytest.mark.parametrize("T,N,K", [(40037, 384, 96), (40037, 96, 96), (70001, 96, 384), (40037, 288, 96)])
def test_gemm_wgrad_long_token_axis(T, N, K):
    """Few output tiles over tens of thousands of tokens: the library picks the reduction split by whole waves of
    the persistent grid (early-stage wgrads); weight and bias gradients accumulate."""
    ops = _ops()
    dy, x = _bf(_rand(T, N, seed=1)), _bf(_rand(T, K, seed=2))
    ref = dy.float().t() @ x.float()
    dw, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    ops.linear_wgrad(dy, x, dw, dbias=db)
    assert rel_err(dw, ref) < 1e-4
    assert rel_err(db, dy.float().sum(0)) < 1e-4
    ops.linear_wgrad(dy, x, dw)
    assert rel_err(dw, 2 * ref) < 1e-4


@pytest.mark.parametrize("M,N,K", [(20000, 384, 96), (20000, 288, 96), (30000, 96, 96), (20000, 96, 384),
                                   (26000, 192, 192), (20000, 96, 288)])
def test_gemm_stationary_resident_weights(M, N, K):
    """Many row tiles and few column tiles: the n-stationary schedule with the weight tile kept resident in the TMA
    ring (fetched on the first pass only; the ring is cut to a multiple of the tile's k-blocks), forward epilogues
    and the MN-major dgrad, with a ragged last row tile."""
    ops = _ops()
    M += 37
    x, w, b = _bf(_rand(M, K, seed=1)), _bf(_rand(N, K, seed=2, scale=K ** -0.5)), _rand(N, seed=3)
    pre = x.float() @ w.float().t() + b
    assert rel_err(ops.linear_fwd(x, w, b).float(), pre) < BF16_TOL
    resid = _rand(M, N, seed=4)
    assert rel_err(ops.linear_fwd(x, w, b, out_dtype=torch.float32, resid=resid), resid + pre) < 1e-5
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    y = ops.linear_fwd(x, w, b, gelu_aux=aux)
    assert rel_err(aux.float(), pre) < BF16_TOL
    assert rel_err(y.float(), torch.nn.functional.gelu(aux.float())) < BF16_TOL
    dy = _bf(_rand(M, K, seed=5))
    w2 = _bf(_rand(K, N, seed=6, scale=N ** -0.5))
    assert rel_err(ops.linear_dgrad(dy, w2).float(), dy.float() @ w2.float()) < BF16_TOL
    dx = ops.linear_dgrad(dy, w2, gelu_aux=aux)
    a = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(a).backward(dy.float() @ w2.float())
    assert rel_err(dx.float(), a.grad) < BF16_TOL


@pytest.mark.parametrize("M,N,K", [(3000, 384, 1536), (1100, 1152, 384), (16128, 384, 384), (257, 768, 192)])
def test_gemm_cta_pairs(M, N, K):
    """Shapes with N, K >= 192 run on CTA pairs (cta_group::2, 256-row tiles): odd numbers of 128-row tiles (the
    peer's half tile lies below the matrix), every store epilogue, K-major and MN-major B operands."""
    ops = _ops()
    x, w, b = _bf(_rand(M, K, seed=1)), _bf(_rand(N, K, seed=2, scale=K ** -0.5)), _rand(N, seed=3)
    pre = x.float() @ w.float().t() + b
    assert rel_err(ops.linear_fwd(x, w, b).float(), pre) < BF16_TOL
    resid = _rand(M, N, seed=4)
    assert rel_err(ops.linear_fwd(x, w, b, out_dtype=torch.float32, resid=resid), resid + pre) < 1e-5
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    y = ops.linear_fwd(x, w, b, gelu_aux=aux)
    assert rel_err(aux.float(), pre) < BF16_TOL
    assert rel_err(y.float(), torch.nn.functional.gelu(aux.float())) < BF16_TOL
    if N % 128 == 0:                       # MN-major B halves are whole 64-column boxes
        dy = _bf(_rand(M, K, seed=5))
        w2 = _bf(_rand(K, N, seed=6, scale=N ** -0.5))
        assert rel_err(ops.linear_dgrad(dy, w2).float(), dy.float() @ w2.float()) < BF16_TOL


@pytest.mark.parametrize("M,N,K", [(512, 96, 96), (3000, 288, 96), (1000, 96, 384), (700, 384, 96),
                                   (5000, 192, 768), (64, 96, 64), (100, 1536, 384)])
def test_gemm_dgrad_wgrad(M, N, K):
    ops = _ops()
    dy, x, w = _bf(_rand(M, N, seed=1)), _bf(_rand(M, K, seed=2)), _bf(_rand(N, K, seed=3, scale=K ** -0.5))
    dx = ops.linear_dgrad(dy, w)
    assert rel_err(dx.float(), dy.float() @ w.float()) < BF16_TOL
    dw = torch.zeros(N, K, device="cuda")
    ops.linear_wgrad(dy, x, dw)
    ref = dy.float().t() @ x.float()
    assert rel_err(dw, ref) < 1e-4
    ops.linear_wgrad(dy, x, dw)           # accumulates
    assert rel_err(dw, 2 * ref) < 1e-4
    # bias gradient from the same GEMM (row sums of dy^T against an all-ones tile), accumulating
    db = torch.ones(N, device="cuda")
    dw2 = torch.zeros_like(dw)
    ops.linear_wgrad(dy, x, dw2, dbias=db)
    assert rel_err(dw2, ref) < 1e-4
    assert rel_err(db - 1, dy.float().sum(0)) < 1e-4


# ----------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,C", [(1000, 96), (333, 768), (50, 3072), (17, 4096), (4000, 192), (70001, 96),
                                    (9, 384), (5000, 1536), (131, 64), (777, 128), (40000, 384)])
def test_layernorm(rows, C):
    ops = _ops()
    x, g, b = _rand(rows, C, seed=1) * 2 + 0.5, 1 + 0.1 * _rand(C, seed=2), 0.1 * _rand(C, seed=3)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, out_dtype=torch.float32)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5)
    assert rel_err(y, ref) < 1e-5
    yb, _, _ = ops.layernorm_fwd(x, g, b)
    assert rel_err(yb.float(), ref) < BF16_TOL
    dy = _rand(rows, C, seed=4)
    ref.backward(dy)
    rg = _rand(rows, C, seed=5)
    dx, dxb = ops.layernorm_bwd(dy, x, mean, rstd, g, resid_grad=rg, want_bf16=True)
    assert rel_err(dx, xr.grad + rg) < 1e-5
    assert rel_err(dxb.float(), xr.grad + rg) < BF16_TOL
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.ln_param_grad(dy, x, mean, rstd, dg, db)
    assert rel_err(dg, gr.grad) < 1e-4 and rel_err(db, br.grad) < 1e-4
    # parameter gradients fused into the backward pass (accumulating)
    dg2, db2 = dg.clone(), db.clone()
    dx2, _ = ops.layernorm_bwd(_bf(dy), x, mean, rstd, g, dgamma=dg2, dbeta=db2)
    assert rel_err(dx2, xr.grad) < BF16_TOL
    gr.grad = None; br.grad = None; xr.grad = None
    torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5).backward(_bf(dy).float())
    assert rel_err(dg2 - dg, gr.grad) < 1e-4 and rel_err(db2 - db, br.grad) < 1e-4
    cs = torch.zeros(C, device="cuda")
    ops.colsum(_bf(dy), cs)
    assert rel_err(cs, _bf(dy).float().sum(0)) < 1e-4


# ----------------------------------------------------------------------------- attention
def _ref_window_attention(qkv, table, B, grid, window, shift, shifted, heads, hd, mask_value=100.0):
    """fp32 torch restatement on the gathered windows (oracle index tables)."""
    T, _ = qkv.shape
    C = heads * hd
    tok = torch.from_numpy(O.window_tokens(grid, window)).cuda()
    nW, N = tok.shape
    src = torch.from_numpy(O.shifted_source(grid, shift)).cuda()[tok] if shifted else tok
    x = qkv.reshape(B, -1, 3, heads, hd)[:, src.reshape(-1)].reshape(B * nW, N, 3, heads, hd)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    rpi = torch.from_numpy(O.relative_position_index(window)).cuda()
    s = s + table[rpi.reshape(-1)].reshape(N, N, heads).permute(2, 0, 1)[None]
    if shifted:
        m = torch.from_numpy(O.shift_mask(grid, window, shift)).cuda() * (mask_value / 100.0)
        s = (s.reshape(B, nW, heads, N, N) + m[None, :, None]).reshape(B * nW, heads, N, N)
    o = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, nW * N, C)
    out = torch.zeros(B, grid[0] * grid[1] * grid[2], C, device="cuda", dtype=o.dtype)
    out[:, src.reshape(-1)] = o
    return out.reshape(T, C)


@pytest.mark.parametrize("grid,heads,shifted", [((12, 14, 12), 2, False), ((12, 14, 12), 2, True),
                                                ((6, 7, 6), 3, True), ((18, 21, 12), 1, True),
                                                # several windows per persistent CTA (ring wrap-around, deferred
                                                # read-outs and MMAs across window boundaries, masked and unmasked mixed)
                                                ((24, 28, 24), 3, True), ((24, 28, 24), 3, False)])
def test_window_attention_fwd_bwd(grid, heads, shifted):
    ops = _ops()
    B, window, shift, hd = 2, (6, 7, 6), (3, 3, 3), 32
    C = heads * hd
    T = B * grid[0] * grid[1] * grid[2]
    qkv = _bf(_rand(T, 3 * C, seed=1))
    table = 0.5 * _rand(11 * 13 * 11, heads, seed=2)
    geom = ops.WindowGeom(B, grid, window, shift if shifted else (0, 0, 0), shifted)
    out, lse = ops.attn_fwd(qkv, heads, hd, S=geom.S, N=geom.N, scale=hd ** -0.5, geom=geom, table=table)
    q32 = qkv.float().requires_grad_(True)
    t32 = table.clone().requires_grad_(True)
    ref = _ref_window_attention(q32, t32, B, grid, window, shift, shifted, heads, hd)
    assert rel_err(out.float(), ref) < BF16_TOL
    dout = _bf(_rand(T, C, seed=3))
    ref.backward(dout.float())
    dtable = torch.zeros_like(table)
    dqkv = ops.attn_bwd(qkv, out, dout, lse, heads, hd, S=geom.S, N=geom.N, scale=hd ** -0.5, geom=geom, table=table,
                        dtable=dtable)
    assert rel_err(dqkv.float(), q32.grad) < 2e-2
    assert rel_err(dtable, t32.grad) < 2e-2


def test_window_attention_finite_mask_leak():
    """The reference's shift mask is a finite -100 (models/swin_transformer_3d.py:486-491): with logits of +-60 nats
    some masked keys leak through it.  The forward applies the mask on the tensor cores as a +100.009-nat offset of
    same-region pairs; rows with a visible leak must still follow the reference's -100 (a 101.8-nat mask -- round 1's
    constant -- is off by more than 10 % on those rows, asserted below so the test stays sensitive)."""
    ops = _ops()
    B, grid, window, shift, heads, hd = 2, (12, 14, 12), (6, 7, 6), (3, 3, 3), 2, 32
    C = heads * hd
    T = B * grid[0] * grid[1] * grid[2]
    x = _rand(T, 3 * C, seed=1)
    x[:, : 2 * C] *= 8.0                                              # q, k: logit std ~ 64 nats
    qkv = _bf(x)
    table = 0.5 * _rand(11 * 13 * 11, heads, seed=2)
    geom = ops.WindowGeom(B, grid, window, shift, True)
    out, lse = ops.attn_fwd(qkv, heads, hd, S=geom.S, N=geom.N, scale=hd ** -0.5, geom=geom, table=table)
    q32 = qkv.float().requires_grad_(True)
    args = (B, grid, window, shift, True, heads, hd)
    ref = _ref_window_attention(q32, table, *args)
    with torch.no_grad():
        hard = _ref_window_attention(qkv.float(), table, *args, mask_value=1e4)      # no leak at all
        old = _ref_window_attention(qkv.float(), table, *args, mask_value=101.82)
    rows = (ref.detach() - hard).abs().amax(1) > 1e-2
    assert int(rows.sum()) > 50
    assert rel_err(old[rows], ref.detach()[rows]) > 5e-2
    assert rel_err(out.float()[rows], ref.detach()[rows]) < 2e-2
    assert rel_err(out.float(), ref.detach()) < BF16_TOL
    dout = _bf(_rand(T, C, seed=3))
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, heads, hd, S=geom.S, N=geom.N, scale=hd ** -0.5, geom=geom, table=table,
                        dtable=torch.zeros_like(table))
    assert rel_err(dqkv.float(), q32.grad) < 3e-2


@pytest.mark.parametrize("N,heads", [(13, 2), (81, 2), (811, 6)])
def test_dense_attention_fwd_bwd(N, heads):
    ops = _ops()
    B, hd = 2, 64
    C = heads * hd
    qkv = _bf(_rand(B * N, 3 * C, seed=1))
    out, lse = ops.attn_fwd(qkv, heads, hd, S=B, N=N, scale=hd ** -0.5)
    q32 = qkv.float().requires_grad_(True)
    x = q32.reshape(B, N, 3, heads, hd)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(B * N, C)
    assert rel_err(out.float(), ref) < BF16_TOL
    dout = _bf(_rand(B * N, C, seed=3))
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, heads, hd, S=B, N=N, scale=hd ** -0.5)
    assert rel_err(dqkv.float(), q32.grad) < 2e-2


# ----------------------------------------------------------------------------- layout kernels (bit-exact)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_patch_gather_bit_exact(dtype):
    ops = _ops()
    B, D, H, W, p = 2, 10, 13, 7, (4, 4, 4)
    vol = _rand(B, 1, D, H, W, seed=1).to(dtype)
    rows = ops.patch_gather(vol, p, out_dtype=torch.float32)
    v = torch.nn.functional.pad(vol.float(), (0, (-W) % 4, 0, (-H) % 4, 0, (-D) % 4))
    gd, gh, gw = v.shape[2] // 4, v.shape[3] // 4, v.shape[4] // 4
    ref = v.reshape(B, gd, 4, gh, 4, gw, 4).permute(0, 1, 3, 5, 2, 4, 6).reshape(B * gd * gh * gw, 64)
    assert torch.equal(rows, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 12, 20, 16), (1, 10, 13, 8), (3, 144, 168, 144)])
def test_patch_gather_444_fast_path_bit_exact(dtype, shape):
    """bf16 rows of 4x4x4 patches with W % 4 == 0 take the vectorised kernel (one 32-byte store per (token, i))."""
    ops = _ops()
    B, D, H, W = shape
    vol = _rand(B, 1, D, H, W, seed=1).to(dtype)
    rows = ops.patch_gather(vol, (4, 4, 4))
    v = torch.nn.functional.pad(vol.float(), (0, 0, 0, (-H) % 4, 0, (-D) % 4))
    gd, gh, gw = v.shape[2] // 4, v.shape[3] // 4, v.shape[4] // 4
    ref = v.reshape(B, gd, 4, gh, 4, gw, 4).permute(0, 1, 3, 5, 2, 4, 6).reshape(B * gd * gh * gw, 64)
    assert rows.dtype == torch.bfloat16 and torch.equal(rows, ref.to(torch.bfloat16))


@pytest.mark.parametrize("dtype,shape,patch", [(torch.float16, (2, 32, 48, 32), (16, 16, 16)),
                                               (torch.float16, (2, 20, 33, 30), (16, 16, 16)),
                                               (torch.float32, (3, 8, 9, 8), (4, 8, 4)),
                                               (torch.bfloat16, (2, 16, 16, 22), (4, 8, 8))])
def test_patch_gather_fused_into_layernorm(dtype, shape, patch):
    """vsn_patch_ln_fwd / vsn_patch_ln_param_grad (ViT Rearrange + LayerNorm(P), models/vit_3d.py:364-371) against the
    composition patch_gather -> layernorm -> column reduction; ragged volumes (zero-padded patches) included."""
    ops = _ops()
    B, D, H, W = shape
    P = patch[0] * patch[1] * patch[2]
    vol = _rand(B, 1, D, H, W, seed=1).to(dtype)
    gamma, beta = 1.0 + 0.1 * _rand(P, seed=2), 0.1 * _rand(P, seed=3)
    rows = ops.patch_gather(vol, patch, out_dtype=torch.float32)
    y_ref, mean_ref, rstd_ref = ops.layernorm_fwd(rows, gamma, beta)
    assert ops.patch_ln_supported(patch, vol)
    y, mean, rstd = ops.patch_ln_fwd(vol, patch, gamma, beta)
    if P > 1024:      # LayerNorm(P) takes the generic kernel, whose lane map and summation order the fused one follows
        assert torch.equal(mean, mean_ref) and torch.equal(rstd, rstd_ref) and torch.equal(y, y_ref)
    else:             # the register-resident LayerNorm kernels sum in another order: fp32 round-off
        assert rel_err(mean, mean_ref) < 1e-5 and rel_err(rstd, rstd_ref) < 1e-5
        assert rel_err(y.float(), y_ref.float()) < 1e-3
    dy = _rand(*y.shape, seed=4).to(y.dtype)
    dg_ref, db_ref = torch.zeros(P, device="cuda"), torch.zeros(P, device="cuda")
    ops.ln_param_grad(dy, rows, mean_ref, rstd_ref, dg_ref, db_ref)
    dg, db = torch.ones(P, device="cuda"), torch.ones(P, device="cuda")               # accumulate (+=)
    ops.patch_ln_param_grad(dy, vol, patch, mean, rstd, dg, db)
    assert rel_err(dg - 1.0, dg_ref) < 1e-4 and rel_err(db - 1.0, db_ref) < 1e-4


def test_grid_copy_and_merge_gather_bit_exact():
    ops = _ops()
    B, C = 2, 32
    real, padded = (7, 8, 5), (12, 14, 6)
    x = _rand(B * real[0] * real[1] * real[2], C, seed=1)
    xp = ops.grid_copy(x, real, padded, B, C)
    ref = torch.nn.functional.pad(x.reshape(B, *real, C), (0, 0, 0, padded[2] - real[2], 0, padded[1] - real[1], 0,
                                                             padded[0] - real[0]))
    assert torch.equal(xp.reshape(B, *padded, C), ref)
    back = ops.grid_copy(xp, padded, real, B, C)
    assert torch.equal(back, x)
    g = ops.merge_gather(xp, padded, real, B, C)
    t = torch.nn.functional.pad(x.reshape(B, *real, C), (0, 0, 0, real[2] % 2, 0, real[1] % 2, 0, real[0] % 2))
    parts = [t[:, a::2, b::2, c::2] for (a, b, c) in
             ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))]
    refg = torch.cat(parts, -1).reshape(-1, 8 * C)
    assert torch.equal(g, refg)
    # scatter is the exact transpose of gather
    dx = ops.merge_scatter(g, padded, real, B, C)
    assert torch.equal(dx.reshape(B, *padded, C)[:, :real[0], :real[1], :real[2]], x.reshape(B, *real, C))
    assert float(dx.abs().sum()) == float(x.abs().sum())


@pytest.mark.parametrize("C,real,padded", [(96, (6, 7, 5), (6, 7, 5)), (96, (7, 8, 5), (12, 14, 6)),
                                           (192, (9, 11, 9), (9, 11, 9)), (192, (4, 3, 6), (6, 7, 6))])
def test_merge_gather_fused_into_layernorm(C, real, padded):
    """vsn_merge_ln_fwd / _bwd (PatchMerging's gather inside its LayerNorm, models/swin_transformer_3d.py:553-572)
    against the two-kernel composition merge_gather + layernorm: the forward bit for bit (same element-to-lane map and
    summation order), the backward to the order of the dgamma / dbeta atomics; odd extents (zero neighbours) and a
    padded stage grid (tokens the backward must leave zero) included."""
    ops = _ops()
    B = 2
    x = _rand(B * real[0] * real[1] * real[2], C, seed=1)
    xp = ops.grid_copy(x, real, padded, B, C) if real != padded else x
    gamma, beta = 1.0 + 0.1 * _rand(8 * C, seed=2), 0.1 * _rand(8 * C, seed=3)
    xg = ops.merge_gather(xp, padded, real, B, C)
    y_ref, mean_ref, rstd_ref = ops.layernorm_fwd(xg, gamma, beta)
    y, mean, rstd = ops.merge_ln_fwd(xp, gamma, beta, padded, real, B, C)
    if C == 96:       # LayerNorm(768) takes the register-resident kernel with the same lane map: bit for bit
        assert torch.equal(y, y_ref) and torch.equal(mean, mean_ref) and torch.equal(rstd, rstd_ref)
    else:             # LayerNorm(1536) takes the generic kernel (another summation order): fp32 round-off
        assert rel_err(mean, mean_ref) < 1e-5 and rel_err(rstd, rstd_ref) < 1e-6
        assert rel_err(y.float(), y_ref.float()) < 1e-3
    dy = _rand(*y.shape, seed=4).to(y.dtype)
    dg_ref, db_ref = torch.zeros(8 * C, device="cuda"), torch.zeros(8 * C, device="cuda")
    dxg, _ = ops.layernorm_bwd(dy, xg, mean_ref, rstd_ref, gamma, dgamma=dg_ref, dbeta=db_ref)
    dx_ref = ops.merge_scatter(dxg, padded, real, B, C)
    dg, db = torch.ones(8 * C, device="cuda"), torch.ones(8 * C, device="cuda")        # accumulate (+=)
    scale = torch.tensor([0.0, 1.0 / 0.7], device="cuda")                              # per-sample DropPath factors
    dx, dxb = ops.merge_ln_bwd(dy, xp, mean, rstd, gamma, dg, db, padded, real, B, C, want_bf16=True, row_scale=scale)
    assert rel_err(dx, dx_ref) < 1e-6
    want_b = (dx.reshape(B, -1) * scale[:, None]).reshape(dx.shape).to(dxb.dtype)
    assert torch.equal(dxb, want_b)
    pad_mask = torch.ones(B, *padded, 1, device="cuda", dtype=torch.bool)
    pad_mask[:, :real[0], :real[1], :real[2]] = False
    assert float((dx.reshape(B, *padded, C) * pad_mask).abs().max()) == 0.0
    assert rel_err(dg - 1.0, dg_ref) < 1e-5 and rel_err(db - 1.0, db_ref) < 1e-5


def test_head_and_token_mean():
    ops = _ops()
    B, T, F_, K = 3, 150, 768, 5
    x, W, b = _rand(B * T, F_, seed=1), _rand(K, F_, seed=2, scale=0.05), _rand(K, seed=3)
    pooled = ops.token_mean(x, B, T, F_)
    assert rel_err(pooled, x.reshape(B, T, F_).mean(1)) < 1e-6
    logits = ops.head_fwd(pooled, W, b)
    assert rel_err(logits, pooled @ W.t() + b) < 1e-5
    dl = _rand(B, K, seed=4)
    dW, db = torch.zeros_like(W), torch.zeros_like(b)
    dfeat = ops.head_bwd(dl, pooled, W, dW, db)
    assert rel_err(dfeat, dl @ W) < 1e-5 and rel_err(dW, dl.t() @ pooled) < 1e-5 and rel_err(db, dl.sum(0)) < 1e-5
    dx = ops.token_mean_bwd(dfeat, B, T, F_)
    assert rel_err(dx.reshape(B, T, F_), (dfeat / T)[:, None].expand(B, T, F_)) < 1e-6
