"""Micro-benchmark of vsn_gemm_bf16 at the hot shapes of the Swin-T step (B = 8).  CUDA-event timing."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vsn_b200  # noqa: E402,F401
from vsn_b200 import ops  # noqa: E402

BF, F32 = torch.bfloat16, torch.float32


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
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    cases = []
    for (M, N, K) in [(435456, 384, 96), (16128, 1536, 384), (54432, 768, 192), (2016, 3072, 768)]:
        x, w, b = r(M, K).to(BF), (0.05 * r(N, K)).to(BF), r(N)
        h = torch.empty(M, N, device="cuda", dtype=BF)
        cases.append((f"fwd+gelu  M{M} N{N} K{K}", lambda x=x, w=w, b=b, h=h: ops.linear_fwd(x, w, b, gelu_aux=h), 2 * M * N * K,
                      2 * M * K + 4 * M * N))
        dy = r(M, N).to(BF)
        w2 = (0.05 * r(K, N)).to(BF)      # fc2 weight [C, 4C]: dgrad gives [M, 4C]
        dyc = r(M, K).to(BF)
        cases.append((f"dgrad+gelu' M{M} N{N} K{K}", lambda dyc=dyc, w2=w2, h=h: ops.linear_dgrad(dyc, w2, gelu_aux=h), 2 * M * N * K,
                      2 * M * K + 4 * M * N))
    for (M, N, K) in [(435456, 288, 96), (435456, 96, 384), (435456, 96, 96), (54432, 192, 768), (54432, 192, 192),
                      (16128, 384, 1536), (16128, 384, 384), (16128, 1152, 384), (2016, 768, 3072)]:
        x, w, b = r(M, K).to(BF), (0.05 * r(N, K)).to(BF), r(N)
        res = r(M, N) if N in (96, 192, 384, 768) else None
        cases.append((f"fwd {'resid f32' if res is not None else 'bf16'} M{M} N{N} K{K}",
                      lambda x=x, w=w, b=b, res=res: ops.linear_fwd(x, w, b, out_dtype=F32 if res is not None else BF, resid=res),
                      2 * M * N * K, 2 * M * K + (8 if res is not None else 2) * M * N))
    for (T, N, K) in [(16128, 384, 384), (16128, 1536, 384), (435456, 384, 96), (435456, 96, 96), (54432, 768, 192)]:
        dy, x = r(T, N).to(BF), r(T, K).to(BF)
        dw = torch.zeros(N, K, device="cuda")
        cases.append((f"wgrad tokens{T} N{N} K{K}", lambda dy=dy, x=x, dw=dw: ops.linear_wgrad(dy, x, dw), 2 * T * N * K,
                      2 * T * (N + K)))
    for name, fn, fl, by in cases:
        if only and only not in name:
            continue
        ms = timeit(fn)
        print(f"{name:42s} {ms:8.4f} ms  {fl / ms / 1e9:7.1f} TF/s  {by / ms / 1e6:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
