"""Build libvsn_b200.so in-tree with nvcc for sm_100a (no torch extension machinery, plain C ABI)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libvsn_b200.so")
LIB_F16 = os.path.join(PKG, "libvsn_b200_f16.so")   # the same sources with -DVSN_F16: IEEE-half operands (precision mode)
SOURCES = ["api.cu", "gemm_tc.cu", "attn.cu", "wattn_tc.cu", "dattn_tc.cu", "norm.cu", "layout.cu", "optim.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math" if False else "-DVSN_B200=1"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), only=None, f16: bool = False) -> str:
    """extra_flags / only (source names) serve measurement builds, e.g. `-DVSN_MBAR_HINT_NS=0` for wattn_tc.cu.
    f16=True builds the half-operand variant (libvsn_b200_f16.so) from the same sources."""
    LIB = LIB_F16 if f16 else globals()["LIB"]
    if not force and not needs_build(LIB):
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build_f16" if f16 else "build")
    if f16:
        extra_flags = [*extra_flags, "-DVSN_F16=1"]
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if only is not None and src not in only and os.path.exists(obj):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-I", HERE, "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    # shared CUDA runtime (the process already holds torch's libcudart.so.12; /usr/local/cuda for stand-alone loads)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    only = [a[len("--only="):] for a in sys.argv[1:] if a.startswith("--only=")] or None
    variants = [True] if "--f16-only" in sys.argv else ([False] if "--bf16-only" in sys.argv else [False, True])
    for f16 in variants:
        print(build(force="--force" in sys.argv or bool(extra), verbose="-v" in sys.argv, extra_flags=extra, only=only, f16=f16))
