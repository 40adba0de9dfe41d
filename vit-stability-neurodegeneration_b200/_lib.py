"""ctypes binding of the C ABI declared in include/vsn_b200.h.

There is no fallback: if the shared library is missing, or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# Precision of the 16-bit operands, fixed for the life of the process (one library, one activation dtype):
#   "bf16" (default)  libvsn_b200.so      bfloat16 operands, fp32 accumulation: logits / gradients within 2e-2
#   "f16"             libvsn_b200_f16.so  the same kernels built with -DVSN_F16: IEEE-half operands, i.e. TF32's
#                     11 significant bits with fp32 accumulation -- the precision mode of north_star's tighter
#                     tolerance and the dtype the reference trains in (autocast(float16) + GradScaler,
#                     train/train_transformer.py:1141-1160); gradients need the usual loss scale.
PRECISION = os.environ.get("VSN_B200_PRECISION", "bf16").lower()
if PRECISION not in ("bf16", "f16"):
    raise RuntimeError(f"VSN_B200_PRECISION must be 'bf16' or 'f16', got {PRECISION!r}")
LIB_PATH = os.path.join(_HERE, "libvsn_b200.so" if PRECISION == "bf16" else "libvsn_b200_f16.so")


def act_dtype():
    """torch dtype of the 16-bit activations / operands this process computes in."""
    import torch
    return torch.bfloat16 if PRECISION == "bf16" else torch.float16


_p, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes (restype is int unless noted)
SIGNATURES = {
    "vsn_version": [],
    "vsn_precision": [],
    "vsn_check_device": [],
    "vsn_gemm_bf16": [_p, _ll, _i, _p, _ll, _i, _i, _i, _i, _p, _ll, _i, _p, _i, _p, _ll, _p, _ll, _p, _i, _f, _i, _p, _p],
    "vsn_layernorm_fwd": [_p, _ll, _p, _p, _p, _ll, _i, _p, _p, _ll, _i, _f, _p],
    "vsn_layernorm_bwd": [_p, _ll, _i, _p, _ll, _p, _p, _p, _p, _ll, _p, _ll, _p, _ll, _p, _i, _p, _p, _ll, _i, _p],
    "vsn_colreduce": [_p, _ll, _i, _p, _ll, _p, _p, _p, _p, _ll, _i, _p],
    "vsn_attn_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _f, _p],
    "vsn_attn_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _f, _p],
    "vsn_patch_gather": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vsn_patch_ln_fwd": [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _f, _p],
    "vsn_patch_ln_param_grad": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p],
    "vsn_grid_copy": [_p, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p],
    "vsn_merge_gather": [_p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _i, _p],
    "vsn_merge_ln_fwd": [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _f, _p],
    "vsn_merge_ln_bwd": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "vsn_mixup_f16": [_p, _p, _p, _p, _i, _ll, _p],
    "vsn_volume_stats_f16": [_p, _p, _p, _i, _ll, _p, _p, _p],
    "vsn_mixup_zscore_f16": [_p, _p, _p, _p, _p, _i, _ll, _p],
    "vsn_tta_views_f16": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "vsn_cast_rows_bf16": [_p, _p, _p, _i, _ll, _i, _p],
    "vsn_cast_bf16": [_p, _p, _ll, _p],
    "vsn_token_mean": [_p, _p, _i, _i, _i, _i, _p],
    "vsn_head_fwd": [_p, _p, _p, _p, _i, _i, _i, _p],
    "vsn_head_bwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "vsn_vit_assemble": [_p, _p, _p, _p, _i, _i, _i, _p],
    "vsn_vit_assemble_bwd": [_p, _p, _p, _i, _i, _i, _p],
    "vsn_mt_chunk_elems": [],
    "vsn_mt_sqnorm": [_p, _p, _p, _p, _p, _i, _p, _i, _p],
    "vsn_sam_scale": [_p, _i, _f, _p, _p, _p],
    "vsn_mt_sam_perturb": [_p, _p, _p, _p, _p, _p, _i, _p, _p, _i, _p],
    "vsn_mt_copy": [_p, _p, _p, _p, _p, _i, _p],
    "vsn_mt_cast_bf16": [_p, _p, _p, _p, _p, _i, _p],
    "vsn_adamw_prepare": [_p, _p, _p, _f, _f, _p, _p],
    "vsn_mt_adamw": [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _f, _f, _f, _f, _f, _i, _p],
    "vsn_mt_ema": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _f, _f, _p],
}

_lib = None
_device_checked = False


def load() -> C.CDLL:
    """Load libvsn_b200.so (building is `__graft_entry__.build()`'s job, not ours)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(vsn_b200 has no fallback path)")
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the binding disagree
            fn.argtypes = args
            fn.restype = C.c_int
        lib.vsn_launch_count.argtypes = []
        lib.vsn_launch_count.restype = C.c_longlong
        lib.vsn_last_error.argtypes = []
        lib.vsn_last_error.restype = C.c_char_p
        if lib.vsn_precision() != (0 if PRECISION == "bf16" else 1):
            raise RuntimeError(f"{LIB_PATH} was not built for VSN_B200_PRECISION={PRECISION}: rebuild it "
                               "(python __graft_entry__.py build)")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().vsn_last_error().decode("utf-8", "replace")


def launch_count() -> int:
    return int(load().vsn_launch_count())


def check_device() -> None:
    global _device_checked
    if not _device_checked:
        lib = load()
        if lib.vsn_check_device() != 0:
            raise RuntimeError("vsn_b200: " + last_error())
        _device_checked = True


# Optional per-call device timing (bench.py / profiling only): when PROFILE is a list, every call is
# bracketed by CUDA events on the current stream and (name, tag, start, end) is appended.
PROFILE = None
TAG = ""
WORK = (0, 0)   # (algorithmic flops, algorithmic bytes) of the next call, set by ops.py when profiling


def call(name: str, *args) -> None:
    lib = load()
    if PROFILE is None:
        rc = getattr(lib, name)(*args)
    else:
        import torch
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = getattr(lib, name)(*args)
        e.record()
        global WORK
        PROFILE.append((name, TAG, WORK, s, e))
        WORK = (0, 0)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {last_error()}")
