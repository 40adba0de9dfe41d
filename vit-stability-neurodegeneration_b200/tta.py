"""Test-time augmentation and snapshot-ensemble inference on the vsn_b200 path.

`TestTimeAugmentation` keeps the constructor and call surface of the reference's class
(eval/test_time_augmentation.py:14-420) so eval/eval_transformer.py:410-461 can drive it.  What changes underneath:
the reference resamples every view on the HOST with MONAI (flip, RandAffine, centre crop + trilinear resize), ships it
to the GPU and runs a batch-1 forward per view -- 8 forwards and 8 H2D copies per subject, times the number of
snapshots.  Every one of those views is an affine map from output voxel to source coordinates, so here the normalised
float16 volume is copied once, ONE kernel writes all views of all subjects of the batch (`ops.tta_views`), the model
runs once on the [B*V] batch, and the entropy-weighted average (:338-354) is a few [V,K] tensor ops on the device.
`SnapshotEnsemble` (scripts/transformer.sh:241-266 averages the predictions of the saved snapshots) re-uses the views
for every snapshot.

The affine views follow MONAI's conventions as far as they can be stated without MONAI in the image: grid centred on
the volume, rotation (about axes 0, 1, 2 in turn) then translation in voxels, bilinear, border padding; their random
parameters come from a numpy RandomState (uniform in +-range), not from MONAI's stream.  Identity, flip and
crop + resize are exact (the latter against torch's `interpolate(mode="trilinear", align_corners=False)`).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops


def _view(mat: np.ndarray, lo: Sequence[float], hi: Sequence[float]) -> np.ndarray:
    return np.concatenate([np.asarray(mat, np.float64).reshape(12), np.asarray(lo, np.float64), np.asarray(hi, np.float64)])


def identity_view(shape) -> np.ndarray:
    return _view(np.hstack([np.eye(3), np.zeros((3, 1))]), (0, 0, 0), [s - 1 for s in shape])


def flip_view(shape, axis: int = 0) -> np.ndarray:
    """RandFlip(prob=1.0, spatial_axis=axis): source index = size-1 - index along the axis."""
    m = np.hstack([np.eye(3), np.zeros((3, 1))])
    m[axis, axis] = -1.0
    m[axis, 3] = shape[axis] - 1
    return _view(m, (0, 0, 0), [s - 1 for s in shape])


def affine_view(shape, rotate: Sequence[float], translate: Sequence[float]) -> np.ndarray:
    """Rotation by `rotate` radians about axes 0, 1, 2 around the volume centre, then `translate` voxels."""
    a, b, c = (float(v) for v in rotate)
    r0 = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    r1 = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
    r2 = np.array([[math.cos(c), -math.sin(c), 0], [math.sin(c), math.cos(c), 0], [0, 0, 1]])
    R = r0 @ r1 @ r2
    ctr = (np.asarray(shape, np.float64) - 1.0) / 2.0
    off = ctr - R @ ctr + np.asarray(translate, np.float64)
    return _view(np.hstack([R, off[:, None]]), (0, 0, 0), [s - 1 for s in shape])


def center_crop_resize_view(shape, scale: float) -> np.ndarray:
    """CenterSpatialCrop(roi = int(size * scale)) then Resize(size, mode="trilinear") (:171-186,296-304): with
    align_corners=False, source = start + (index + 0.5) * roi/size - 0.5, clamped to the crop box."""
    m = np.zeros((3, 4))
    lo, hi = [], []
    for ax, n in enumerate(shape):
        roi = int(n * scale)
        start = max(n // 2 - roi // 2, 0)
        m[ax, ax] = roi / n
        m[ax, 3] = start + 0.5 * roi / n - 0.5
        lo.append(start)
        hi.append(start + roi - 1)
    return _view(m, lo, hi)


class TestTimeAugmentation:
    """eval/test_time_augmentation.py:14-420, batched on the device.  `model` is a vsn_b200 model (eval mode is the
    caller's business, as in the reference); inputs are the NORMALISED volumes [B,C=1,D,H,W] (or [C,D,H,W]) on the host
    or the device, any float dtype."""

    __test__ = False      # not a pytest class

    def __init__(self, model: torch.nn.Module, device: torch.device, num_samples: int = 5, use_flip: bool = True,
                 use_affine: bool = True, use_scaled_center_crop: bool = True, crop_roi_scale: float = 0.9,
                 affine_rotate_range: Tuple[float, float, float] = (3.0, 3.0, 3.0),
                 affine_translate_range: Tuple[float, float, float] = (5.0, 5.0, 5.0),
                 target_shape: Optional[Tuple[int, int, int]] = None, use_amp: bool = False,
                 use_channels_last: bool = True, use_entropy_weighting: bool = True, seed: Optional[int] = None):
        self.model = model
        self.device = torch.device(device)
        self.num_samples = num_samples
        self.use_flip = use_flip
        self.use_affine = use_affine
        self.use_scaled_center_crop = use_scaled_center_crop
        self.crop_roi_scale = crop_roi_scale
        if target_shape is not None:
            raise NotImplementedError("vsn_b200 TTA resizes the crop back to the input shape (target_shape=None)")
        self.target_shape = None
        self.use_amp = use_amp                      # the path computes in bf16 either way
        self.use_channels_last = use_channels_last  # a layout hint of the reference; C = 1 here
        self.use_entropy_weighting = use_entropy_weighting
        self.rotate_range = tuple(math.radians(r) for r in affine_rotate_range)
        self.translate_range = tuple(float(t) for t in affine_translate_range)
        self._rng = np.random.RandomState(seed)

    # -- views ---------------------------------------------------------------------------------------
    def build_views(self, shape) -> np.ndarray:
        """[V, 18] in the reference's order (:246-309): identity, flip, num_samples affine draws, centre crop."""
        shape = tuple(int(s) for s in shape)
        views = [identity_view(shape)]
        if self.use_flip:
            views.append(flip_view(shape, 0))
        if self.use_affine:
            for _ in range(self.num_samples):
                rot = [self._rng.uniform(-r, r) for r in self.rotate_range]
                tr = [self._rng.uniform(-t, t) for t in self.translate_range]
                views.append(affine_view(shape, rot, tr))
        if self.use_scaled_center_crop:
            views.append(center_crop_resize_view(shape, self.crop_roi_scale))
        return np.stack(views)

    # -- averaging (:338-354) --------------------------------------------------------------------------
    def combine(self, probs: torch.Tensor) -> torch.Tensor:
        """probs [B, V, K] -> [B, K]: inverse-entropy weights w = 1/(H + 1e-6), H = -sum p log p with p clamped at
        1e-10 (:199-219), normalised over the views; uniform mean otherwise."""
        if probs.shape[1] == 1:
            return probs[:, 0]
        if not self.use_entropy_weighting:
            return probs.mean(dim=1)
        p = probs.clamp(min=1e-10)
        ent = -(p * p.log()).sum(-1)
        wts = 1.0 / (ent + 1e-6)
        wts = wts / wts.sum(dim=1, keepdim=True)
        return (probs * wts[..., None]).sum(dim=1)

    @torch.no_grad()
    def view_batch(self, x: torch.Tensor) -> Tuple[torch.Tensor, int]:
        if x.ndim == 4:
            x = x[None]
        x = x.to(self.device, non_blocking=True)
        if x.dtype != torch.float16:
            x = x.half()
        views = self.build_views(x.shape[2:])
        mats = torch.from_numpy(views.astype(np.float32)).to(self.device)
        return ops.tta_views(x.contiguous(), mats), views.shape[0]

    @torch.no_grad()
    def predict(self, x: torch.Tensor) -> torch.Tensor:
        """Averaged softmax probabilities [B, K] (or [K] for an unbatched [C,D,H,W] input), on the device."""
        single = x.ndim == 4
        xv, V = self.view_batch(x)
        probs = torch.softmax(self.model(xv).float(), dim=1)
        out = self.combine(probs.view(-1, V, probs.shape[-1]))
        return out[0] if single else out

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.predict(x)

    def get_num_augmentations(self) -> int:
        return 1 + int(self.use_flip) + (self.num_samples if self.use_affine else 0) + int(self.use_scaled_center_crop)


class SnapshotEnsemble:
    """Mean of the (TTA) predictions of several snapshots of one architecture (scripts/transformer.sh:241-266).  The
    views of a batch are generated once; each snapshot's weights are loaded into the one resident model."""

    def __init__(self, model: torch.nn.Module, snapshots: Sequence[Dict[str, torch.Tensor]], tta: Optional[TestTimeAugmentation] = None):
        self.model, self.snapshots, self.tta = model, list(snapshots), tta
        if not self.snapshots:
            raise ValueError("SnapshotEnsemble needs at least one state_dict")

    @torch.no_grad()
    def predict(self, x: torch.Tensor) -> torch.Tensor:
        dev = next(self.model.parameters()).device
        if self.tta is not None:
            xv, V = self.tta.view_batch(x)
        else:
            xv, V = (x if x.ndim == 5 else x[None]).to(dev), 1
        total = None
        for sd in self.snapshots:
            self.model.load_state_dict(sd)
            probs = torch.softmax(self.model(xv).float(), dim=1)
            p = self.tta.combine(probs.view(-1, V, probs.shape[-1])) if self.tta is not None else probs
            total = p if total is None else total + p
        return total / len(self.snapshots)

    __call__ = predict


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n subjects for `rank`: sizes differ by at most one, the first n % world ranks take
    the longer ones (a rank may get nothing when n < world)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedInference:
    """BASELINE config 5 on several GPUs: inference is embarrassingly parallel over subjects, so every
    rank predicts one contiguous shard of the subject list with its own predictor (a `TestTimeAugmentation`, a
    `SnapshotEnsemble` or the bare model wrapped in softmax -- all views and snapshots of a subject stay on the rank
    that owns it) and ONE all-gather of the `[n, K]` probabilities at the end gives every rank the full table in subject
    order.  No collective inside the data path.  The reference evaluates checkpoints one process at a time
    (eval/eval_transformer.py:1052-1157, scripts/transformer.sh:241-266); this is the batch-sharded form of that loop.

    `predictor(x[b,1,D,H,W]) -> [b, K]` probabilities; `group=None` is the default process group; without an
    initialised process group the class degrades to the plain loop (world 1)."""

    def __init__(self, predictor, group=None, batch: int = 2):
        import torch.distributed as dist
        self.predictor, self.group, self.batch = predictor, group, max(1, int(batch))
        self._dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self._dist.get_world_size(group) if self._dist else 1
        self.rank = self._dist.get_rank(group) if self._dist else 0

    @torch.no_grad()
    def predict_local(self, volumes: torch.Tensor) -> Tuple[torch.Tensor, Tuple[int, int]]:
        """This rank's shard of `volumes` [N,1,D,H,W] (host or device): ([hi - lo, K] probabilities, (lo, hi))."""
        lo, hi = shard_range(volumes.shape[0], self.world, self.rank)
        outs = [self.predictor(volumes[i:min(i + self.batch, hi)]).float() for i in range(lo, hi, self.batch)]
        return (torch.cat(outs) if outs else torch.zeros(0, 0)), (lo, hi)

    @torch.no_grad()
    def gather(self, local: torch.Tensor, n: int, device=None) -> torch.Tensor:
        """All ranks call this once: `[n, K]` in subject order on every rank (ranks without subjects join with zero rows)."""
        if self.world == 1:
            return local
        dist = self._dist
        if device is None and not local.numel() and dist.get_backend(self.group) == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())      # a rank without subjects still needs a CUDA buffer
        dev = local.device if local.numel() else (torch.device(device) if device is not None else local.device)
        k = torch.tensor([local.shape[1] if local.ndim == 2 else 0], device=dev, dtype=torch.int64)
        dist.all_reduce(k, op=dist.ReduceOp.MAX, group=self.group)          # empty shards do not know K
        K = int(k.item())
        cap = -(-n // self.world)                                            # longest shard
        pad = torch.zeros(cap, K, device=dev, dtype=torch.float32)
        if local.numel():
            pad[:local.shape[0]] = local.to(dev)
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        rows = []
        for r, part in enumerate(parts):
            lo, hi = shard_range(n, self.world, r)
            rows.append(part[:hi - lo])
        return torch.cat(rows)

    def predict(self, volumes: torch.Tensor, device=None) -> torch.Tensor:
        local, _ = self.predict_local(volumes)
        return self.gather(local, volumes.shape[0], device=device)

    __call__ = predict
