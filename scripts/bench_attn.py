"""Micro-benchmark of the window-attention kernels at the Swin-T stage shapes (B = 8 volumes).
Prints ms per call and algorithmic TFLOP/s (fwd 4*S*N^2*C, bwd 8*S*N^2*C) with CUDA events."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vsn_b200  # noqa: E402,F401
from vsn_b200 import ops  # noqa: E402

STAGES = [((36, 42, 36), 3), ((18, 21, 18), 6), ((12, 14, 12), 12), ((6, 7, 6), 24)]


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B, window, shift, hd = 8, (6, 7, 6), (3, 3, 3), 32
    do_bwd = "--fwd-only" not in sys.argv
    for grid, heads in STAGES:
        for shifted in (False, True):
            C = heads * hd
            T = B * grid[0] * grid[1] * grid[2]
            g = torch.Generator(device="cuda").manual_seed(0)
            qkv = torch.randn(T, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
            table = 0.1 * torch.randn(11 * 13 * 11, heads, device="cuda", generator=g)
            geom = ops.WindowGeom(B, grid, window, shift if shifted else (0, 0, 0), shifted)
            kw = dict(S=geom.S, N=geom.N, scale=hd ** -0.5, geom=geom, table=table)
            out, lse = ops.attn_fwd(qkv, heads, hd, **kw)
            ms = timeit(lambda: ops.attn_fwd(qkv, heads, hd, **kw))
            fl = 4 * geom.S * geom.N * geom.N * C
            line = f"grid {grid} heads {heads:2d} shifted {int(shifted)}: fwd {ms:.4f} ms {fl / ms / 1e9:7.1f} TF/s"
            if do_bwd:
                dout = torch.randn(T, C, device="cuda", generator=g).to(torch.bfloat16)
                dtable = torch.zeros_like(table)
                msb = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, heads, hd, dtable=dtable, **kw))
                line += f" | bwd {msb:.4f} ms {2 * fl / msb / 1e9:7.1f} TF/s"
            print(line, flush=True)


if __name__ == "__main__":
    main()
