"""Golden vectors for the input step (TEST INFRASTRUCTURE): the UNMODIFIED reference's `MRIMixUp`
(/root/reference/dataset/dataset.py:186-286, seeded branch) over a small synthetic float16 cohort, followed by monai's
`NormalizeIntensity()` -- restated here, monai is not installed: (x - mean) / std over the whole image in float32,
population std, no division when std == 0 (train/train_transformer.py:1729-1752 puts it last in the transform chain).

    python oracle/make_golden_input.py        # writes tests/golden/mixup.npz (needs /root/reference)
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


class NormalizeIntensity:
    def __call__(self, x):
        x = x.float()
        s = x.std(unbiased=False)
        return (x - x.mean()) / s if float(s) != 0.0 else x - x.mean()


class Cohort(torch.utils.data.Dataset):
    def __init__(self, vols, labels, diagnoses):
        self.vols, self.labels = vols, labels
        self.meta_data = pd.DataFrame({"Subject": [f"s{i}" for i in range(len(diagnoses))], "Diagnosis": diagnoses})
        self.trace = []

    def __getitem__(self, idx):
        self.trace.append(int(idx))
        return self.vols[idx].clone(), self.labels[idx].clone()

    def __len__(self):
        return len(self.vols)


def main():
    refshim.install()
    # dataset/__init__.py also imports the NIfTI preprocessing (nibabel, nilearn, monai's CenterSpatialCrop): absent here
    # and not on this path -- empty stand-ins so that the package imports
    import types
    for name, attrs in (("nibabel", ()), ("nilearn", ()), ("nilearn.image", ("resample_img",))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, None)
            sys.modules[name] = m
    mt = sys.modules["monai.transforms"]
    if not hasattr(mt, "CenterSpatialCrop"):
        mt.CenterSpatialCrop = type("CenterSpatialCrop", (), {})
    from dataset.dataset import MRIMixUp
    g = torch.Generator().manual_seed(11)
    N, K, shape = 12, 3, (1, 8, 12, 8)
    diagnoses = [["AD", "CN", "FTD"][i % K] for i in range(N)]
    vols = (torch.rand(N, *shape, generator=g) * 900.0 + 40.0 * torch.randn(N, *shape, generator=g)).half()
    vols[5] = 3.0                                                # a constant volume: std == 0
    labels = torch.zeros(N, K)
    for i, d in enumerate(diagnoses):
        labels[i, ["AD", "CN", "FTD"].index(d)] = 1.0
    base = Cohort(vols, labels, diagnoses)
    out = dict(volumes=vols.numpy(), labels=labels.numpy(), diagnoses=np.array([["AD", "CN", "FTD"].index(d) for d in diagnoses]))
    for epoch in (0, 3):
        mix = MRIMixUp(base, num_samples=N, alpha=0.3, mixup_prob=0.7, transform=NormalizeIntensity(), seed=7)
        mix._current_epoch = epoch
        xs, ys, partners = [], [], []
        for idx in range(N):
            base.trace.clear()
            x, y = mix[idx]
            xs.append(x.numpy())
            ys.append(y.numpy())
            partners.append(base.trace[1] if len(base.trace) > 1 else -1)
        out[f"x_e{epoch}"] = np.stack(xs).astype(np.float32)
        out[f"y_e{epoch}"] = np.stack(ys).astype(np.float32)
        out[f"partner_e{epoch}"] = np.array(partners, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "mixup.npz"), **out)
    print("mixup.npz:", {k: v.shape for k, v in out.items()}, "partners e0", out["partner_e0"].tolist())


if __name__ == "__main__":
    main()
