"""Split-K sweep of the late-stage weight-gradient GEMMs at the fused step's token counts (16 volumes per pass):
dW[N,K] += dY[T,N]^T X[T,K] with the bias gradient riding along, library-chosen split (0) against fixed splits.
Measurement script (CUDA events, L2-sized rotation of the operands so that every launch streams them from HBM/L2 as in
the step)."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vsn_b200  # noqa: E402,F401
from vsn_b200 import ops  # noqa: E402

BF = ops.BF16


def timeit(fn, iters=20):
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
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(32256, 1536, 384), (32256, 384, 1536), (32256, 1152, 384), (32256, 384, 384),
              (108864, 768, 192), (108864, 192, 768), (108864, 576, 192), (108864, 192, 192),
              (4032, 3072, 768), (4032, 768, 3072), (4032, 2304, 768), (4032, 768, 768)]
    splits = [0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48]
    for T, N, K in shapes:
        dy = torch.randn(T, N, device="cuda", generator=g).to(BF)
        x = torch.randn(T, K, device="cuda", generator=g).to(BF)
        dw = torch.zeros(N, K, device="cuda")
        db = torch.zeros(N, device="cuda")
        res = []
        for s in splits:
            fn = lambda s=s: ops.gemm(dy, dy.stride(0), 1, x, x.stride(0), 1, N, K, T, dw, K, 2, split_k=s, rowsum=db)  # noqa: E731
            try:
                res.append((s, timeit(fn)))
            except RuntimeError as e:
                res.append((s, float("nan")))
        auto = res[0][1]
        best = min(res[1:], key=lambda r: r[1] if r[1] == r[1] else 1e9)
        print(f"tokens{T} N{N} K{K}: auto {auto * 1e3:.1f} us | best split {best[0]} {best[1] * 1e3:.1f} us ({auto / best[1]:.2f}x) | " +
              " ".join(f"{s}:{ms * 1e3:.1f}" for s, ms in res[1:]), flush=True)


if __name__ == "__main__":
    main()
