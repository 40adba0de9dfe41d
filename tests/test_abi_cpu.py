"""The C-ABI shared library loads without a GPU and exports every symbol that include/vsn_b200.h declares; the
ctypes binding covers exactly that set.  No compute call is made here."""
import ctypes
import os
import re

import vsn_b200  # noqa: F401
from vsn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "vsn_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(vsn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = header_symbols()
    for must in ("vsn_gemm_bf16", "vsn_attn_fwd", "vsn_attn_bwd", "vsn_layernorm_fwd", "vsn_layernorm_bwd",
                 "vsn_patch_gather", "vsn_merge_gather", "vsn_mt_sam_perturb", "vsn_mt_ema", "vsn_last_error"):
        assert must in syms


def test_library_exports_every_header_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python __graft_entry__.py build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_matches_header():
    bound = set(_lib.SIGNATURES) | {"vsn_last_error", "vsn_launch_count"}
    assert bound == set(header_symbols())
    lib = _lib.load()
    assert lib.vsn_version() >= 100


def test_precision_build_exports_the_same_abi():
    """libvsn_b200_f16.so (-DVSN_F16: IEEE-half operands) is the same ABI; each build names its own encoding."""
    f16 = os.path.join(os.path.dirname(_lib.LIB_PATH), "libvsn_b200_f16.so")
    assert os.path.exists(f16), "build the library first: python __graft_entry__.py build"
    lib = ctypes.CDLL(f16)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.vsn_precision() == 1 and _lib.load().vsn_precision() == (0 if _lib.PRECISION == "bf16" else 1)
