"""vsn_gemm_bf16 against cuBLAS (torch F.linear) on large and stage-2 shapes.  The GPU is kept busy with a spin
kernel while the timed launches are enqueued, so the figures are kernel time, not host enqueue time."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vsn_b200  # noqa: E402,F401
from vsn_b200 import ops  # noqa: E402

BF = torch.bfloat16


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(4_000_000)          # ~2 ms: the launches below queue up behind it
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


SHAPES = [(8192, 8192, 8192), (8192, 4096, 4096), (16128, 1536, 384), (16128, 384, 1536), (16128, 1152, 384),
          (16128, 384, 384), (16128, 1536, 1536), (16128, 256, 1536), (16128, 512, 1536), (2016, 3072, 768),
          (2016, 768, 3072)]
for (M, N, K) in SHAPES:
    x = torch.randn(M, K, device="cuda").to(BF)
    w = (0.05 * torch.randn(N, K, device="cuda")).to(BF)
    ms = timeit(lambda: ops.linear_fwd(x, w, None))
    ms2 = timeit(lambda: torch.nn.functional.linear(x, w))
    print(f"M{M} N{N} K{K}: ours {ms * 1e3:8.1f} us {2 * M * N * K / ms / 1e9:7.1f} TF/s | cublas {ms2 * 1e3:8.1f} us "
          f"{2 * M * N * K / ms2 / 1e9:7.1f} TF/s", flush=True)
