"""Portable synthetic weights / inputs shared by the golden generator and the tests.

numpy's MT19937 `RandomState` is bit-stable across platforms and versions, so a
fixture only has to store outputs: weights and inputs are re-derived from
(key, shape, seed) on whichever box the test runs.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import zlib
from typing import Dict, Sequence, Tuple

import numpy as np


def _rs(tag: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(tag.encode()) ^ (seed * 2654435761)) & 0xFFFFFFFF)


def synth_param(key: str, shape: Sequence[int], seed: int = 0) -> np.ndarray:
    """Deterministic fp32 tensor for a state_dict key.  Scales are chosen so every
    term of the path matters numerically (non-trivial biases, LN gains, bias table)."""
    r = _rs(key, seed).standard_normal(tuple(shape)).astype(np.float32)
    leaf = key.rsplit(".", 1)[-1]
    if "relative_position_bias_table" in key:
        return (0.5 * r).astype(np.float32)
    if key in ("pos_embedding", "cls_token"):
        return (0.3 * r).astype(np.float32)
    is_norm = (".norm" in key or "norm." in key or key.startswith("to_patch_embedding.1")
               or key.startswith("to_patch_embedding.3") or key.startswith("mlp_head.0")
               or ".net.0." in key)
    if is_norm:
        return (1.0 + 0.1 * r if leaf == "weight" else 0.1 * r).astype(np.float32)
    if leaf == "bias":
        return (0.05 * r).astype(np.float32)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
    return (r / np.sqrt(fan_in)).astype(np.float32)


def synth_state(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, np.ndarray]:
    return {k: synth_param(k, s, seed) for k, s in shapes.items()}


def synth_volume(shape: Sequence[int], seed: int = 0) -> np.ndarray:
    """fp32 volume holding fp16-representable values, MNI-like: gaussian inside a
    centred ellipsoid, zero outside, z-scored."""
    B = shape[0]
    D, H, W = shape[-3:]
    v = _rs("volume", seed).standard_normal(tuple(shape)).astype(np.float32)
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, D), np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing="ij")
    inside = ((zz / 0.9) ** 2 + (yy / 0.9) ** 2 + (xx / 0.9) ** 2) <= 1.0
    v = v * inside.astype(np.float32)
    v = v.reshape(B, -1)
    v = (v - v.mean(1, keepdims=True)) / (v.std(1, keepdims=True) + 1e-8)
    return v.reshape(tuple(shape)).astype(np.float16).astype(np.float32)


def synth_targets(batch: int, classes: int, seed: int = 0) -> np.ndarray:
    """Soft targets: one-hot for even samples, a MixUp-style 0.7/0.3 blend for odd ones."""
    rs = _rs("targets", seed)
    t = np.zeros((batch, classes), dtype=np.float32)
    for b in range(batch):
        i = int(rs.randint(classes))
        if b % 2 == 0:
            t[b, i] = 1.0
        else:
            j = (i + 1 + int(rs.randint(classes - 1))) % classes
            t[b, i], t[b, j] = 0.7, 0.3
    return t


def synth_keep_masks(n_calls: int, batch: int, keep: float = 0.7, seed: int = 0) -> np.ndarray:
    """[n_calls, B] 0/1 keep decisions for DropPath, with at least one 0 and one 1 per call when B>1."""
    rs = _rs("droppath", seed)
    m = (rs.uniform(size=(n_calls, batch)) < keep).astype(np.float32)
    if batch > 1:
        m[:, 0] = 1.0
        m[0::2, 1] = 0.0
    return m
