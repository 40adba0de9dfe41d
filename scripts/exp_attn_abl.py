"""GPU experiment: window-attention backward at the stage-0 shape (B = 8) under the ablation switches of a
-DVSN_ABL build (environment VSN_ABL=<mask>, see csrc/wattn_tc.cu); prints ms per call."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vsn_b200  # noqa: E402,F401
from vsn_b200 import ops  # noqa: E402
from bench_attn import timeit  # noqa: E402


def main():
    B, window, shift, hd = 8, (6, 7, 6), (3, 3, 3), 32
    out_line = f"VSN_ABL={os.environ.get('VSN_ABL', '0'):>4}:"
    for grid, heads in (((36, 42, 36), 3), ((12, 14, 12), 12)):
        for shifted in (False, True):
            C = heads * hd
            T = B * grid[0] * grid[1] * grid[2]
            g = torch.Generator(device="cuda").manual_seed(0)
            qkv = torch.randn(T, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
            table = 0.1 * torch.randn(11 * 13 * 11, heads, device="cuda", generator=g)
            geom = ops.WindowGeom(B, grid, window, shift if shifted else (0, 0, 0), shifted)
            kw = dict(S=geom.S, N=geom.N, scale=hd ** -0.5, geom=geom, table=table)
            out, lse = ops.attn_fwd(qkv, heads, hd, **kw)
            dout = torch.randn(T, C, device="cuda", generator=g).to(torch.bfloat16)
            dtable = torch.zeros_like(table)
            ms = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, heads, hd, dtable=dtable, **kw), iters=20)
            out_line += f"  S{geom.S} h{heads} m{int(shifted)} {ms:.4f} ms"
    print(out_line, flush=True)


if __name__ == "__main__":
    main()
