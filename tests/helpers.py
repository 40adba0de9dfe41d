"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def meta():
    with open(os.path.join(GOLD, "meta.json")) as f:
        return json.load(f)


def golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def rel_err(a, b):
    a = torch.as_tensor(a).double().reshape(-1).cpu()
    b = torch.as_tensor(b).double().reshape(-1).cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def synth_sd(shapes, dtype=torch.float32, device="cpu", seed=0):
    from oracle.synth import synth_state
    fl = {k: tuple(v) for k, v in shapes.items() if "relative_position_index" not in k}
    return {k: torch.from_numpy(v).to(device=device, dtype=dtype) for k, v in synth_state(fl, seed).items()}
