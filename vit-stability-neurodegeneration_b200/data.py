"""The input step before the path, on the device.

The reference keeps a float16 `.pt` cache of the volumes, mixes pairs on the host inside a DataLoader worker
(`dataset/dataset.py:186-286`, `MRIMixUp`), normalises every volume with monai's `NormalizeIntensity()` on the host
(`train/train_transformer.py:1729-1752`) and ships float32 to the GPU (`:1122-1128`).  Here the host only decides WHO is
mixed with WHOM and how much -- `mixup_plan` restates the seeded branch of `MRIMixUp.__getitem__` decision for decision --
and ships the raw float16 volumes; MixUp, the whole-image statistics and the normalisation run on the device in two
passes over the batch (`ops.mixup_zscore`), and the volumes stay float16 until the patch gather.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

_MAX_UINT32 = 2 ** 32      # utils/seed.py


def mixup_plan(idx: int, *, seed: int, epoch: int, diagnoses: Sequence, class_indices: Dict[object, Sequence[int]],
               class_list: Sequence, alpha: float, mixup_prob: float) -> Tuple[Optional[int], float]:
    """dataset/dataset.py:232-268, seeded branch: returns (partner index or None, weight of sample `idx`).

    RandomState(seed + epoch + idx): one uniform draw decides whether the sample is left alone, then the partner's
    class among the OTHER classes, then the partner inside that class, then Beta(alpha, alpha)."""
    rng = np.random.RandomState(int((seed + epoch + idx) % _MAX_UINT32))
    if bool(rng.rand() > mixup_prob):
        return None, 1.0
    cls1 = diagnoses[idx]
    available = [c for c in class_list if c != cls1]
    cls2 = available[int(rng.randint(0, len(available)))]
    members = class_indices[cls2]
    idx2 = int(members[int(rng.randint(0, len(members)))])
    return idx2, float(rng.beta(alpha, alpha))


class DeviceInputPipeline:
    """Batches of raw float16 volumes -> mixed, z-scored float16 volumes and soft labels on the device.

    `volumes` [N,1,D,H,W] float16 (the reference's `.pt` cache, pinned host memory or already resident), `labels` [N,K]
    one-hot float32, `diagnoses` the class of every subject.  `batch(indices, epoch)` copies the samples and their MixUp
    partners (each volume once), mixes and normalises on the device and returns `(x, y)` for `model(x)` /
    `train.TrainStep.step`.  `mixup_prob = 0` is the validation path (normalisation only)."""

    def __init__(self, volumes: torch.Tensor, labels: torch.Tensor, diagnoses: Sequence, *, alpha: float = 0.3,
                 mixup_prob: float = 0.0, seed: int = 0, device: Optional[torch.device] = None):
        if volumes.dtype != torch.float16 or volumes.ndim != 5:
            raise ValueError("DeviceInputPipeline expects the float16 [N,1,D,H,W] volume cache")
        self.volumes, self.labels, self.diagnoses = volumes, labels, list(diagnoses)
        self.alpha, self.mixup_prob, self.seed = float(alpha), float(mixup_prob), int(seed)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        # dataset/dataset.py:213-217: indices grouped by class, classes in groupby (sorted) order
        self.class_list = sorted(set(self.diagnoses))
        self.class_indices = {c: [i for i, d in enumerate(self.diagnoses) if d == c] for c in self.class_list}

    def plan(self, indices: Sequence[int], epoch: int) -> List[Tuple[Optional[int], float]]:
        if self.mixup_prob <= 0.0:
            return [(None, 1.0)] * len(indices)
        return [mixup_plan(int(i), seed=self.seed, epoch=epoch, diagnoses=self.diagnoses,
                           class_indices=self.class_indices, class_list=self.class_list, alpha=self.alpha,
                           mixup_prob=self.mixup_prob) for i in indices]

    def batch(self, indices: Sequence[int], epoch: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        plan = self.plan(indices, epoch)
        B = len(indices)
        # rows of the device batch: the B samples first, then every partner that is not among them (each volume once)
        rows = {int(i): r for r, i in enumerate(indices)} if len(set(indices)) == B else {}
        src = [int(i) for i in indices]
        perm, lam = [], []
        for r, (partner, a) in enumerate(plan):
            if partner is None:
                perm.append(r)
                lam.append(1.0)
                continue
            if partner not in rows:
                rows[partner] = len(src)
                src.append(partner)
            perm.append(rows[partner])
            lam.append(a)
        sel = torch.as_tensor(src, dtype=torch.long)
        x = self.volumes.index_select(0, sel.to(self.volumes.device)).to(self.device, non_blocking=True)
        y = self.labels.index_select(0, sel.to(self.labels.device)).to(self.device, non_blocking=True).float()
        mixing = any(p is not None for p, _ in plan)
        lam_t = torch.tensor(lam, dtype=torch.float32).to(self.device, non_blocking=True) if mixing else None
        perm_t = torch.tensor(perm, dtype=torch.int32).to(self.device, non_blocking=True) if mixing else None
        if len(src) > B and mixing:            # partners sit behind the batch: statistics / output for the first B only
            lam_t = torch.cat([lam_t, torch.ones(len(src) - B, device=self.device)])
            perm_t = torch.cat([perm_t, torch.arange(B, len(src), device=self.device, dtype=torch.int32)])
        xn, _ = ops.mixup_zscore(x.contiguous(), lam_t, perm_t)
        if mixing:
            # dataset/dataset.py:281: target1.mul_(alpha).add_(target2, alpha=1 - alpha)
            y = lam_t[:, None] * y + (1.0 - lam_t[:, None]) * y[perm_t.long()]
        return xn[:B], y[:B]
