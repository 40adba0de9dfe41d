"""GPU experiment: run-to-run reproducibility of the attention kernels on identical inputs (relative L2 difference of
repeated calls; 0 = bit-identical).  fp32 atomics may reorder sums (~1e-7); anything near 1e-3 is a race."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vsn_b200  # noqa: E402,F401
from vsn_b200 import ops  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    # dense (ViT): S sequences of N tokens, head_dim 64
    for S, N, heads in ((2, 811, 6), (24, 811, 6), (4, 300, 2)):
        hd, C = 64, heads * 64
        qkv = torch.randn(S * N, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
        dout = torch.randn(S * N, C, device="cuda", generator=g).to(torch.bfloat16)
        outs, grads = [], []
        for _ in range(6):
            out, lse = ops.attn_fwd(qkv, heads, hd, S=S, N=N, scale=hd ** -0.5)
            grads.append(ops.attn_bwd(qkv, out, dout, lse, heads, hd, S=S, N=N, scale=hd ** -0.5).clone())
            outs.append(out.clone())
        C3 = grads[0].shape[1] // 3
        print(f"dense S{S} N{N} h{heads}: fwd max diff {max(rel(o, outs[0]) for o in outs[1:]):.2e} | "
              f"dQ {max(rel(x[:, :C3], grads[0][:, :C3]) for x in grads[1:]):.2e} "
              f"dK {max(rel(x[:, C3:2 * C3], grads[0][:, C3:2 * C3]) for x in grads[1:]):.2e} "
              f"dV {max(rel(x[:, 2 * C3:], grads[0][:, 2 * C3:]) for x in grads[1:]):.2e}", flush=True)
    # window (Swin)
    B, window, shift, hd = 8, (6, 7, 6), (3, 3, 3), 32
    for grid, heads in (((36, 42, 36), 3), ((12, 14, 12), 12)):
        for shifted in (False, True):
            C = heads * hd
            T = B * grid[0] * grid[1] * grid[2]
            qkv = torch.randn(T, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
            dout = torch.randn(T, C, device="cuda", generator=g).to(torch.bfloat16)
            table = 0.1 * torch.randn(11 * 13 * 11, heads, device="cuda", generator=g)
            geom = ops.WindowGeom(B, grid, window, shift if shifted else (0, 0, 0), shifted)
            kw = dict(S=geom.S, N=geom.N, scale=hd ** -0.5, geom=geom, table=table)
            outs, grads, tabs = [], [], []
            for _ in range(5):
                out, lse = ops.attn_fwd(qkv, heads, hd, **kw)
                dt = torch.zeros_like(table)
                grads.append(ops.attn_bwd(qkv, out, dout, lse, heads, hd, dtable=dt, **kw).clone())
                outs.append(out.clone())
                tabs.append(dt)
            print(f"window grid{grid} h{heads} shifted{int(shifted)}: fwd {max(rel(o, outs[0]) for o in outs[1:]):.2e} | "
                  f"dqkv {max(rel(x, grads[0]) for x in grads[1:]):.2e} dtable {max(rel(x, tabs[0]) for x in tabs[1:]):.2e}",
                  flush=True)


if __name__ == "__main__":
    main()
