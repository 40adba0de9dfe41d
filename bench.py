#!/usr/bin/env python
"""Headline benchmark: Swin-3D training volumes/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the UNMODIFIED reference on host cores
    python bench.py --impl reference-cuda                          # the unmodified reference, eager, on the B200
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # N > 1, one rank per GPU over NCCL

A "step" is one optimiser step of the reference's hot loop (train/train_transformer.py:1104-1298) on synthetic
MNI-shaped fp16 volumes: `--micro-batches` accumulation micro-batches of `--batch` volumes per GPU (default
2 x 8 = the reference's per-GPU schedule at 8 GPUs, EFFECTIVE_BATCH_SIZE 128), soft-target CE, AdamW(fused) and
the EMA update (BASELINE config 2); `--sam --classes 5 --mixup` is BASELINE config 3 (SAM two-pass step, MixUp on
the device); `--model vit` is the ViT-3D vit-3c configuration (B = 24).  Scaling is weak (per-GPU work fixed).

  value      whole-job volumes/s with the inputs already resident in HBM
  e2e        same loop through the public API (`vsn_b200.train.TrainStep`) with pinned HOST inputs: H2D copies of
             every micro-batch and the D2H read of the loss are inside the timed region
  roofline   the kernel FAMILY with the largest share of the step (per-launch CUDA-event times from an instrumented
             eager pass of the same step, run on every rank) against MEASURED_PEAKS.json, plus per-family tables
  cpu_baseline  the unmodified reference (baseline/_ref, `kind: reference`) or, when it is absent, the oracle port,
             timed on the host cores
  extras     the north_star target configuration (swin-5c, SAM + EMA + MixUp) measured in the same run, and the
             unmodified reference run eagerly on the same B200 (fp16 autocast + TF32) when baseline/_ref exists
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SWIN = dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=[6, 7, 6], patch_size=[4, 4, 4])
VIT = dict(embed_dim=384, depth=12, num_heads=6, patch_size=(16, 16, 16), img_size=(144, 160, 144), mlp_ratio=4.0)
VOLUMES = {"swin": (144, 168, 144), "vit": (144, 160, 144)}
FLOP_FWD_PER_VOL = {"swin": 108.73e9, "vit": 49.11e9}    # forward on the grids the reference runs


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-cuda"])
    ap.add_argument("--model", default="swin", choices=["swin", "vit"])
    ap.add_argument("--precision", default=None, choices=["bf16", "f16"],
                    help="16-bit operand encoding of the kernels (default: VSN_B200_PRECISION or bf16); f16 = the "
                         "IEEE-half build, stepped through a GradScaler as the reference's fp16 loop")
    ap.add_argument("--batch", type=int, default=None, help="volumes per micro-batch per GPU (BATCH_SIZE: 8 / 24)")
    ap.add_argument("--micro-batches", type=int, default=2, help="accumulation micro-batches per step per GPU")
    ap.add_argument("--classes", type=int, default=3)
    ap.add_argument("--sam", action="store_true", help="SAM two-pass step (BASELINE config 3)")
    ap.add_argument("--mixup", action="store_true", help="MixUp the volumes and labels on the device every step")
    ap.add_argument("--no-ema", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of graph replay")
    ap.add_argument("--no-fuse-micro", action="store_true",
                    help="accumulate the micro-batches one forward/backward at a time, as the reference's loop does, "
                         "instead of one pass over their concatenation (same gradients)")
    ap.add_argument("--graph-comm", action="store_true",
                    help="N>1, graph mode: capture the bucket all-reduces INSIDE the last micro-batch's graph (overlapped "
                         "with backward) instead of reducing after the last replay.  Measured slower on 2 B200 (1672 vs "
                         "1704 volumes/s): the NCCL kernels take SMs from the persistent one-CTA-per-SM grids they overlap")
    ap.add_argument("--no-graph-comm", action="store_true", help="(the default since round 2; kept for old command lines)")
    ap.add_argument("--torch-ddp", action="store_true", help="N>1: use torch DDP instead of vsn_b200.ddp.GradAllReduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the instrumented pass (roofline = null)")
    ap.add_argument("--no-extras", action="store_true", help="skip the swin-5c SAM line and the eager-CUDA reference")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel table of the instrumented pass here")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--stub", action="store_true", help=argparse.SUPPRESS)   # tests: control flow on CPU / gloo
    args = ap.parse_args(argv)
    if args.batch is None:
        args.batch = 8 if args.model == "swin" else 24
    return args


# ------------------------------------------------------------------------------------------- synthetic data
def synth_batch(batch, classes, seed, volume, mixup=True):
    """fp16 [B,1,D,H,W] MNI-like volumes + soft labels, generated on the host."""
    import torch
    g = torch.Generator().manual_seed(seed)
    D, H, W = volume
    v = torch.randn(batch, 1, D, H, W, generator=g)
    zz = torch.linspace(-1, 1, D).view(D, 1, 1)
    yy = torch.linspace(-1, 1, H).view(1, H, 1)
    xx = torch.linspace(-1, 1, W).view(1, 1, W)
    inside = ((zz / 0.9) ** 2 + (yy / 0.9) ** 2 + (xx / 0.9) ** 2) <= 1.0
    v = v * inside
    flat = v.view(batch, -1)
    v = ((flat - flat.mean(1, keepdim=True)) / (flat.std(1, keepdim=True) + 1e-8)).view_as(v)
    y = torch.zeros(batch, classes)
    cls = torch.randint(classes, (batch,), generator=g)
    y[torch.arange(batch), cls] = 1.0
    if mixup:                                    # MixUp-style soft targets on half of the samples
        lam = torch.distributions.Beta(0.3, 0.3).sample((batch,)).clamp(0.05, 0.95)
        other = (cls + 1 + torch.randint(classes - 1, (batch,), generator=g)) % classes
        mix = torch.rand(batch, generator=g) < 0.5
        for b in range(batch):
            if mix[b]:
                y[b] = 0
                y[b, cls[b]] = lam[b]
                y[b, other[b]] = 1 - lam[b]
    return v.half(), y


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- reference arms
def _ref_ctor(args):
    import torch
    if args.model == "swin":
        return dict(in_channels=1, mlp_ratio=4.0, qkv_bias=True, dropout=0.0, attention_dropout=0.0,
                    stochastic_depth_prob=0.15, num_classes=args.classes, norm_layer=torch.nn.LayerNorm, **SWIN)
    return dict(num_classes=args.classes, in_channels=1, dropout=0.0, attention_dropout=0.0, **VIT)


def _reference_modules():
    """The UNMODIFIED reference (baseline/_ref, or /root/reference in the build container) through oracle/refshim."""
    from oracle import refshim
    root = refshim.install()
    from models.swin_transformer_3d import SwinTransformerT
    from models.vit_3d import ViTS
    from regularization.sam import SAM
    from regularization.label_smoothing import LabelSmoothingLoss
    from utils.ema import EMAModel
    return dict(root=root, swin=SwinTransformerT, vit=ViTS, SAM=SAM, EMA=EMAModel, loss=LabelSmoothingLoss,
                uninstall=refshim.uninstall)


def _reference_step_fn(args, R, device, batch, micro_batches, amp):
    """One optimiser step of the reference's loop (train/train_transformer.py:1104-1298) on its own modules:
    micro-batch accumulation, LabelSmoothingLoss, AdamW(fused on CUDA) or SAM(AdamW), EMAModel.update."""
    import torch
    torch.manual_seed(0)
    model = R[args.model](**_ref_ctor(args)).to(device).train()
    groups = [{"params": [p for n, p in model.named_parameters() if not (n.endswith(".bias") or p.ndim == 1)]},
              {"params": [p for n, p in model.named_parameters() if n.endswith(".bias") or p.ndim == 1],
               "weight_decay": 0.0}]
    kw = dict(lr=1e-4, weight_decay=0.05)
    if device.type == "cuda":
        kw["fused"] = True
    opt = R["SAM"](groups, torch.optim.AdamW, rho=0.05, adaptive=False, **kw) if args.sam else torch.optim.AdamW(groups, **kw)
    ema = None if args.no_ema else R["EMA"](model=model, decay=0.999, device=device)
    crit = R["loss"](smoothing=0.1)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=100) if amp else None
    data = [synth_batch(batch, args.classes, seed=99 + i, volume=VOLUMES[args.model]) for i in range(micro_batches)]
    data = [(x.float().to(device), y.to(device)) for x, y in data]

    def passes():
        for x, y in data:
            with torch.autocast("cuda", enabled=bool(amp)):
                loss = crit(model(x), y) / len(data)
            (scaler.scale(loss) if scaler else loss).backward()

    def step():
        passes()
        if args.sam:
            if scaler:
                scaler.unscale_(opt.base_optimizer)
            opt.first_step(zero_grad=True)
            passes()
            if scaler:
                opt.second_step(zero_grad=True, scaler=scaler)
            else:
                opt.second_step(zero_grad=True)
        else:
            if scaler:
                scaler.step(opt)
                scaler.update()
            else:
                opt.step()
            opt.zero_grad(set_to_none=True)
        if ema is not None:
            ema.update(model)
    return step


def cpu_arm(args, steps, warmup):
    """The reference's own CPU path on the host cores (all threads): `kind: reference` when baseline/_ref (or the
    checkout) is there, else the oracle port (`kind: port`).  One step = fwd + bwd + optimiser + EMA over
    `--cpu-batch` volumes; returns (cpu_baseline dict, seconds per step, the config it actually ran)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vol = VOLUMES[args.model]
    ran = {"workload": f"{args.model}-{args.classes}c {'SAM(AdamW)' if args.sam else 'AdamW'}"
                       f"{'' if args.no_ema else '+EMA'} training step, CPU fp32",
           "volume": list(vol), "micro_batch": args.cpu_batch, "micro_batches_per_step": 1,
           "global_batch": args.cpu_batch, "parallelism": "cpu"}
    try:
        R = _reference_modules()
        kind, what = "reference", f"unmodified reference modules from {R['root']}"
        step = _reference_step_fn(args, R, torch.device("cpu"), args.cpu_batch, 1, amp=False)
    except (RuntimeError, ImportError) as e:
        if args.model != "swin":
            raise
        kind, what = "port", f"oracle/swin3d_oracle.py (reference not available: {e})"
        step = _oracle_step_fn(args)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    if kind == "reference":
        R["uninstall"]()
    sec = sum(times) / len(times)
    return ({"value": args.cpu_batch / sec, "unit": "volumes/s", "cores": cores, "kind": kind,
             "sample": f"{len(times)} optimiser steps of {args.cpu_batch} volumes {vol} after {warmup} warm-up "
                       f"(fwd+bwd+{'SAM(AdamW)' if args.sam else 'AdamW'}{'' if args.no_ema else '+EMA'}), fp32, "
                       f"{what}, {sec:.2f} s/step"}, sec, ran)


def _oracle_step_fn(args):
    import torch
    from oracle import swin3d_oracle as O
    torch.manual_seed(0)
    sd = {}
    for k, s in swin_state_shapes(args.classes).items():
        leaf = k.rsplit(".", 1)[-1]
        if "norm" in k and leaf == "weight":
            t = torch.ones(s)
        elif leaf == "bias":
            t = torch.zeros(s)
        else:
            t = torch.nn.init.trunc_normal_(torch.empty(s), std=0.02)
        sd[k] = t.requires_grad_(True)
    opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, weight_decay=0.05)
    state = {"snaps": [[v.detach().numpy().copy() for v in sd.values()]]}
    x16, y = synth_batch(args.cpu_batch, args.classes, seed=99, volume=VOLUMES["swin"])
    x = x16.float()

    def step():
        opt.zero_grad(set_to_none=True)
        z = O.swin_forward(sd, x, patch=SWIN["patch_size"], window=SWIN["window_size"], depths=SWIN["depths"],
                           heads=SWIN["num_heads"])
        O.soft_target_ce(z, y, 0.1).backward()
        opt.step()
        state["snaps"] = (state["snaps"] + [[v.detach().numpy().copy() for v in sd.values()]])[-3:]
        _ = [O.ema_average([s[i] for s in state["snaps"]], 0.999) for i in range(len(state["snaps"][0]))]
    return step


def swin_state_shapes(classes):
    """state_dict shapes of Swin-T at the benchmark geometry (parameter container built on the host)."""
    import torch
    import vsn_b200  # noqa: F401
    from vsn_b200.swin_model import SwinTransformerT
    m = SwinTransformerT(in_channels=1, mlp_ratio=4.0, qkv_bias=True, dropout=0.0, attention_dropout=0.0,
                         stochastic_depth_prob=0.0, num_classes=classes, norm_layer=torch.nn.LayerNorm, **SWIN)
    return {k: tuple(v.shape) for k, v in m.state_dict().items() if v.is_floating_point()}



def tta_ensemble_arm(args, dev, subjects=2, snapshots=10, steps=3, warmup=1, be=None):
    """BASELINE config 5: Swin-3D (swin-5c) evaluation with test-time augmentation (8 views: identity, flip, 5 affine,
    centre crop + resize) and a 10-snapshot ensemble (eval/test_time_augmentation.py:221-354,
    scripts/transformer.sh:241-266).  A step = `subjects` normalised fp16 volumes PER GPU copied from pinned host
    memory, all views written by one kernel, one [subjects*8] forward per snapshot (weights swapped in with
    load_state_dict), the entropy-weighted averages and the D2H read of the probabilities.  On N > 1 GPUs
    (`be.world`; every rank calls this) the subject list of a step is N * subjects long, every rank predicts its
    contiguous shard (vsn_b200.tta.ShardedInference) and one all-gather hands every rank the full [N*subjects, K]
    table: no collective in the data path, `scaling: weak`.  volumes/s = all subjects of a step / step time (max over
    ranks)."""
    import torch
    import vsn_b200  # noqa: F401
    from vsn_b200.swin_model import SwinTransformerT
    from vsn_b200.tta import TestTimeAugmentation, SnapshotEnsemble, ShardedInference
    world = be.world if be is not None else 1
    ok, err = 1, ""
    try:
        torch.manual_seed(0)
        model = SwinTransformerT(in_channels=1, mlp_ratio=4.0, qkv_bias=True, dropout=0.0, attention_dropout=0.0,
                                 stochastic_depth_prob=0.15, num_classes=5, norm_layer=torch.nn.LayerNorm, **SWIN).to(dev).eval()
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        snaps = [sd0] + [{k: (v + 1e-3 * torch.randn_like(v) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
                         for _ in range(snapshots - 1)]
        tta = TestTimeAugmentation(model, dev, num_samples=5, seed=0)
        ens = SnapshotEnsemble(model, snaps, tta)
        vol = VOLUMES["swin"]
        host = torch.randn(subjects * world, 1, *vol, generator=torch.Generator().manual_seed(77)).half().pin_memory()
        sharded = ShardedInference(ens, batch=subjects)
    except RuntimeError as e:                # e.g. out of memory on one rank: every rank must learn of it before a collective
        ok, err = 0, str(e).splitlines()[0][:200]
    if world > 1:
        t = torch.tensor([ok], device=dev)
        be.dist.all_reduce(t, op=be.dist.ReduceOp.MIN)
        ok = int(t.item())
    if not ok:
        return {"unavailable": err or "another rank failed to set the arm up"}
    probs = [None]

    def step():
        probs[0] = sharded.predict(host, device=dev).cpu()

    if be is not None:
        for _ in range(warmup):
            step()
        ms = be.timed(step, steps) / steps
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                e0.record()
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    p = probs[0]
    views = tta.get_num_augmentations()
    n = subjects * world
    return {"value": round(n / (ms * 1e-3), 2), "unit": "volumes/s", "ms_per_step": round(ms, 3),
            "config": {"workload": "Swin-3D (Swin-T, swin-5c) inference, test-time augmentation + snapshot ensemble",
                       "volume": list(vol), "subjects_per_step": n, "subjects_per_gpu": subjects, "n_gpus": world,
                       "views": views, "snapshots": snapshots, "forward_volumes_per_step": n * views * snapshots,
                       "parallelism": f"subjects sharded over {world} GPU(s), one all-gather of [{n}, 5] per step" if world > 1 else "1 GPU"},
            "forward_volumes_per_s": round(n * views * snapshots / (ms * 1e-3), 1),
            "h2d_bytes_per_step": int(host.numel() * 2 // world), "d2h_bytes_per_step": int(p.numel() * 4),
            "probabilities_sum_to_one": bool(p.shape[0] == n and float((p.sum(-1) - 1).abs().max()) < 1e-4)}


def eager_cuda_arm(args, steps=3, warmup=2):
    """The unmodified reference, eager PyTorch on this GPU: fp16 autocast + GradScaler + TF32 (the trainer's own
    precision setup, train/train_transformer.py:91-92,1068-1072,1141), same batch schedule as the native arm."""
    import torch
    R = _reference_modules()
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        dev = torch.device("cuda", torch.cuda.current_device())
        step = _reference_step_fn(args, R, dev, args.batch, args.micro_batches, amp=True)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": round(args.batch * args.micro_batches / (ms / 1e3), 2), "unit": "volumes/s",
                "ms_per_step": round(ms, 2), "steps": steps, "warmup": warmup,
                "how": f"unmodified reference modules ({R['root']}), eager PyTorch on the same GPU, fp16 autocast + "
                       f"GradScaler + TF32, {args.micro_batches} x {args.batch} volumes per step, inputs resident",
                "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    finally:
        R["uninstall"]()


def config_dict(args, world):
    vol = VOLUMES[args.model]
    name = f"Swin-3D (Swin-T, swin-{args.classes}c geometry)" if args.model == "swin" else f"ViT-3D (ViT-S, vit-{args.classes}c geometry)"
    in_mb = args.batch * args.micro_batches * vol[0] * vol[1] * vol[2] * 2 / 1e6
    return {"workload": f"{name} {'SAM(AdamW)' if args.sam else 'AdamW'}{'' if args.no_ema else '+EMA'}"
                        f"{'+MixUp+z-score on device' if args.mixup else ''} training step",
            "volume": list(vol), "micro_batch": args.batch, "micro_batches_per_step": args.micro_batches,
            "global_batch": args.batch * args.micro_batches * world, "parallelism": f"dp{world}",
            "micro_batches_fused": bool(args.impl == "native" and not args.no_fuse_micro and not args.torch_ddp
                                        and args.micro_batches > 1),
            "l2": f"per-step inputs ({in_mb:.0f} MB) and activations (GBs) exceed the 126 MB L2; no explicit flush"}


def reference_main(args):
    """`--impl reference`: rank 0 alone times the reference's CPU path; other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.impl == "reference-cuda":
        import torch
        torch.cuda.set_device(0)
        r = eager_cuda_arm(args, steps=max(1, args.steps), warmup=max(1, args.warmup))
        print(json.dumps({"impl": "reference-cuda", "metric": f"{args.model}3d_train_volumes_per_s", **r,
                          "n_gpus": 1, "higher_is_better": True, "dtype": "f16-autocast", "data": "synthetic",
                          "config": config_dict(args, 1)}))
        return
    cb, sec, ran = cpu_arm(args, max(1, args.steps), max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": f"{args.model}3d_train_volumes_per_s", "value": cb["value"], "unit": "volumes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, 1), "config_ran": ran, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- roofline tables
FAMILIES = {"vsn_gemm_bf16": "gemm", "vsn_attn_fwd": "attention_fwd", "vsn_attn_bwd": "attention_bwd",
            "vsn_layernorm_fwd": "layernorm", "vsn_layernorm_bwd": "layernorm", "vsn_colreduce": "layernorm",
            "vsn_merge_ln_fwd": "layernorm", "vsn_merge_ln_bwd": "layernorm", "vsn_patch_ln_fwd": "layernorm",
            "vsn_patch_ln_param_grad": "layernorm"}


def summarise_profile(prof, step_ms, peaks):
    """(name, tag, (flops, bytes), ms) records of one instrumented step -> per-call table + roofline object."""
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = ("measured (MEASURED_PEAKS.json, sustained bf16 / copy bandwidth)" if peaks
                else "fallback (B200_PROFILING.md)")
    groups = {}
    for name, tag, work, ms in prof:
        g = groups.setdefault((name, tag), {"n": 0, "ms": 0.0, "flops": work[0], "bytes": work[1]})
        g["n"] += 1
        g["ms"] += ms
    tot = sum(g["ms"] for g in groups.values()) or 1e-9
    table, fam = [], {}
    for (name, tag), g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
        per = g["ms"] / g["n"]
        row = {"call": name, "shape": tag, "launches": g["n"], "ms_total": round(g["ms"], 4),
               "ms_per_call": round(per, 5), "share": round(g["ms"] / tot, 4)}
        if g["flops"]:
            row["tflops"] = round(g["flops"] / (per * 1e-3) / 1e12, 2)
        if g["bytes"]:
            row["gbs"] = round(g["bytes"] / (per * 1e-3) / 1e9, 1)
        # time this call would take at its own roofline: max(flops / tensor peak, bytes / HBM peak)
        ideal = max(g["flops"] / (tc_peak * 1e12), g["bytes"] / (hbm_peak * 1e9)) * 1e3
        row["frac_of_own_roofline"] = round(ideal / per, 4) if per > 0 and ideal > 0 else None
        table.append(row)
        f = fam.setdefault(FAMILIES.get(name, "elementwise_layout_optim"),
                           {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0, "ideal_ms": 0.0})
        f["ms"] += g["ms"]
        f["launches"] += g["n"]
        f["flops"] += g["flops"] * g["n"]
        f["bytes"] += g["bytes"] * g["n"]
        f["ideal_ms"] += ideal * g["n"]
    families = {}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        families[k] = {"ms": round(f["ms"], 3), "share": round(f["ms"] / tot, 4), "launches": f["launches"],
                       "tflops": round(f["flops"] / (f["ms"] * 1e-3) / 1e12, 1) if f["flops"] else None,
                       "gbs": round(f["bytes"] / (f["ms"] * 1e-3) / 1e9, 1) if f["bytes"] else None,
                       "frac_of_own_roofline": round(f["ideal_ms"] / f["ms"], 4) if f["ms"] else None}
    top_name = next(iter(families))
    top = fam[top_name]
    tensor_bound = top["flops"] / (tc_peak * 1e12) >= top["bytes"] / (hbm_peak * 1e9)
    if tensor_bound:
        ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": round(ach, 2), "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": round(ach / tc_peak, 4), "traffic": None}
    else:
        ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(ach / hbm_peak, 4), "traffic": None}
    roofline["algorithmic_bytes_per_step"] = int(top["bytes"])
    roofline.update({"kernel": top_name, "what": "kernel family with the largest share of the step: algorithmic "
                     "flops (bytes) of all its launches / their summed CUDA-event time",
                     "launches_per_step": top["launches"], "ms_per_step": round(top["ms"], 3),
                     "share_of_step": round(top["ms"] / step_ms, 4), "frac_of_own_roofline": families[top_name]["frac_of_own_roofline"],
                     "peak_source": peak_src, "families": families, "instrumented_step_ms": round(step_ms, 3)})
    # the metric's second half: window-attention TFLOP/s vs peak (heaviest shape per direction)
    wa = {}
    for r in table:
        if r["call"] in ("vsn_attn_fwd", "vsn_attn_bwd") and "tflops" in r:
            key = r["call"].replace("vsn_attn_", "") + (":shifted" if "mask=1" in r["shape"] else ":plain")
            if key not in wa or r["ms_total"] > wa[key]["ms_total"]:
                wa[key] = {"shape": r["shape"], "tflops": r["tflops"], "frac_of_peak": round(r["tflops"] / tc_peak, 4),
                           "ms_per_call": r["ms_per_call"], "ms_total": r["ms_total"]}
    roofline["window_attention"] = wa
    # HBM-bound kernels of the path (LayerNorm, casts, gathers, SAM/EMA): achieved GB/s of the heaviest shape
    hb = {}
    for r in table:
        if r["call"] not in ("vsn_gemm_bf16", "vsn_attn_fwd", "vsn_attn_bwd") and "gbs" in r:
            if r["call"] not in hb or r["ms_total"] > hb[r["call"]]["ms_total"]:
                hb[r["call"]] = {"shape": r["shape"], "gbs": r["gbs"], "frac_of_hbm_peak": round(r["gbs"] / hbm_peak, 3),
                                 "ms_per_call": r["ms_per_call"], "ms_total": r["ms_total"]}
    roofline["hbm_kernels"] = hb
    return table, roofline, tot


# ------------------------------------------------------------------------------------------- backends
class CudaBackend:
    """The real thing: the vsn_b200 model + TrainStep on this rank's GPU."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the vsn_b200 path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        import vsn_b200  # noqa: F401
        from vsn_b200 import _lib
        self._lib = _lib

    # -- workload ------------------------------------------------------------------------------
    def build(self, args):
        torch = self.torch
        from vsn_b200.train import TrainStep
        torch.manual_seed(0)
        if args.model == "swin":
            from vsn_b200.swin_model import SwinTransformerT
            model = SwinTransformerT(in_channels=1, mlp_ratio=4.0, qkv_bias=True, dropout=0.0, attention_dropout=0.0,
                                     stochastic_depth_prob=0.15, num_classes=args.classes,
                                     norm_layer=torch.nn.LayerNorm, **SWIN).to(self.dev)
        else:
            from vsn_b200.vit_model import ViTS
            model = ViTS(num_classes=args.classes, in_channels=1, dropout=0.0, attention_dropout=0.0, **VIT).to(self.dev)
        model.train()
        ddp, sync = None, None
        if self.world > 1:
            if args.torch_ddp:
                ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[self.local], output_device=self.local,
                                                                gradient_as_bucket_view=True)
            else:
                from vsn_b200.ddp import GradAllReduce
                sync = GradAllReduce(model.parameters(), bucket_mb=25.0, buffers=model.buffers())
        graph = not args.no_graph and ddp is None
        scaler = None
        if self._lib.PRECISION == "f16":      # half-precision gradients need the reference's loss scaling (:1141-1160)
            scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=100)
        ts = TrainStep(model, use_sam=args.sam, use_ema=not args.no_ema, ddp_model=ddp, grad_sync=sync, graph=graph,
                       graph_comm=args.graph_comm and not args.no_graph_comm, fuse_micro_batches=not args.no_fuse_micro, scaler=scaler)
        G = args.micro_batches
        vol = VOLUMES[args.model]
        host = [synth_batch(args.batch, args.classes, seed=1234 + self.rank * 100 + i, volume=vol) for i in range(G)]
        host = [(x.pin_memory(), y.pin_memory()) for x, y in host]
        resident = [(x.to(self.dev), y.to(self.dev)) for x, y in host]
        return Workload(self, args, model, ts, host, resident)

    # -- plumbing ------------------------------------------------------------------------------
    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K calls of fn bracketed by barrier + synchronize, timed with CUDA events; max over ranks (ms)."""
        torch = self.torch
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
            self.dist.barrier()
        return ms

    def launch_count(self):
        return self._lib.launch_count()

    def teardown(self):
        if self.world > 1:
            self.sync_all()                  # nobody tears the group down while another rank still communicates
            self.dist.destroy_process_group()


class Workload:
    def __init__(self, be, args, model, ts, host, resident):
        torch = be.torch
        self.be, self.args, self.model, self.ts, self.host, self.resident = be, args, model, ts, host, resident
        self.h2d_bytes = sum(x.numel() * x.element_size() + y.numel() * y.element_size() for x, y in host)
        self.vols_per_step = args.batch * args.micro_batches * be.world
        # e2e staging: two sets of device buffers; the H2D copies of step i+1 are enqueued on a copy stream right
        # after step i's kernels, so they overlap its compute; every timed step still contains exactly one H2D of
        # all its inputs (the copy for the first timed step is issued by the preceding warm-up step, the last timed
        # step issues one for a step that is never run) and the D2H read of its loss.
        self.copy_stream = torch.cuda.Stream()
        self.bufs = [[[torch.empty_like(x, device=be.dev), torch.empty_like(y, device=be.dev)] for x, y in host]
                     for _ in range(2)]
        self.evs = [[torch.cuda.Event() for _ in host] for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.state = {"cur": 0, "primed": False}
        self.mix = None
        if args.mixup:
            # MixUp decisions for 64 steps, drawn once on the host (lambda ~ Beta(0.3, 0.3) with probability 0.5 per
            # sample, else 1; partner = another sample of the micro-batch), resident on the device: a step only
            # indexes the tables, so there is no host work or synchronisation inside the timed region
            g = torch.Generator().manual_seed(77 + be.rank)
            n, G, B = 64, len(host), args.batch
            lam = torch.distributions.Beta(0.3, 0.3).sample((n, G, B))
            lam = torch.where(torch.rand(n, G, B, generator=g) < 0.5, lam, torch.ones_like(lam)).float()
            shift = torch.randint(1, max(B, 2), (n, G, 1), generator=g)
            perm = ((torch.arange(B).view(1, 1, B) + shift) % B).int()
            self.mix = {"lam": lam.to(be.dev), "perm": perm.to(be.dev), "i": 0, "n": n}

    def _prepare(self, batches):
        """BASELINE config 3: MixUp (dataset/dataset.py:230-286) and the z-score that follows it in the reference's
        transform chain (train/train_transformer.py:1729-1752) on the device -- input preparation, three kernels per
        micro-batch for the volumes; the [B, K] labels are mixed by torch."""
        if self.mix is None:
            return batches
        from vsn_b200 import ops
        mx = self.mix
        i = mx["i"] % mx["n"]
        mx["i"] += 1
        out = []
        for j, (x, y) in enumerate(batches):
            lam, perm = mx["lam"][i, j], mx["perm"][i, j]
            # MixUp + NormalizeIntensity() of the raw fp16 volumes in two passes (statistics, mix + z-score + write)
            out.append((ops.mixup_zscore(x, lam, perm)[0], lam[:, None] * y + (1 - lam[:, None]) * y[perm.long()]))
        return out

    def step_resident(self):
        return self.ts.step(self._prepare(self.resident))

    def _enqueue_h2d(self, slot):
        torch = self.be.torch
        self.copy_stream.wait_event(self.done[slot])      # the step that last read this staging set is finished
        with torch.cuda.stream(self.copy_stream):
            for (hx, hy), (dx, dy), ev in zip(self.host, self.bufs[slot], self.evs[slot]):
                dx.copy_(hx, non_blocking=True)
                dy.copy_(hy, non_blocking=True)
                ev.record(self.copy_stream)

    def step_e2e(self):
        torch = self.be.torch
        main_stream = torch.cuda.current_stream()
        st = self.state
        cur = st["cur"]
        if not st["primed"]:
            self.done[0].record(main_stream)
            self.done[1].record(main_stream)
            self._enqueue_h2d(cur)
            st["primed"] = True
        batches = []
        for (dx, dy), ev in zip(self.bufs[cur], self.evs[cur]):
            main_stream.wait_event(ev)
            batches.append((dx, dy))
        loss = self.ts.step(self._prepare(batches))
        self.done[cur].record(main_stream)
        self._enqueue_h2d(cur ^ 1)                        # next step's inputs travel while this step computes
        st["cur"] = cur ^ 1
        return float(loss.item())                         # D2H read of the step's result

    def launches_per_step(self, eager_delta_per_step):
        ts = self.ts
        passes = 2 if self.args.sam else 1
        replays = 1 if ts.fuse_micro_batches else self.args.micro_batches
        graph = ts.graph_kernel_nodes * replays * passes if ts.use_graph else 0
        return int(eager_delta_per_step + graph)

    def replicas_in_sync(self):
        """N > 1: every rank saw different volumes, so identical weights after all these steps mean the gradient
        exchange delivered the same mean gradient everywhere (bit for bit on a fp64 checksum of all parameters)."""
        torch, dist, be = self.be.torch, self.be.dist, self.be
        if be.world == 1:
            return None
        cks = torch.stack([p.detach().double().sum() for p in self.model.parameters()]).sum().reshape(1)
        allc = [torch.empty_like(cks) for _ in range(be.world)]
        dist.all_gather(allc, cks)
        ok = bool(all(torch.equal(c, allc[0]) for c in allc))
        if not ok:
            raise RuntimeError(f"replicas diverged: parameter checksums {[float(c) for c in allc]}")
        return ok

    def instrumented_step(self):
        """One eager (non-graph) step with every C-ABI call bracketed by CUDA events.  Runs on EVERY rank: the step
        contains the gradient exchange."""
        torch, lib, ts = self.be.torch, self.be._lib, self.ts
        was = ts.use_graph
        ts.use_graph = False
        try:
            self.step_resident()                          # eager warm-up (allocator)
            lib.PROFILE = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.step_resident()
            e1.record()
            torch.cuda.synchronize()
            prof, lib.PROFILE = lib.PROFILE, None
        finally:
            lib.PROFILE = None
            ts.use_graph = was
        return [(n, t, w, s.elapsed_time(e)) for n, t, w, s, e in prof], e0.elapsed_time(e1)


class StubBackend:
    """tests/test_bench_cpu.py: the same control flow on CPU tensors over gloo, with a step that contains a
    collective (as the real step contains the gradient all-reduce) -- a rank-asymmetric call deadlocks the test."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = 0
        if self.world > 1:
            dist.init_process_group("gloo")
        self._launches = 0

    def build(self, args):
        return StubWorkload(self, args)

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()

    def timed(self, fn, steps):
        self.sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ms = (time.perf_counter() - t0) * 1e3
        if self.world > 1:
            t = self.torch.tensor([ms])
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
            self.dist.barrier()
        return ms

    def launch_count(self):
        return self._launches

    def teardown(self):
        if self.world > 1:
            self.sync_all()
            self.dist.destroy_process_group()


class StubWorkload:
    def __init__(self, be, args):
        self.be, self.args = be, args
        self.h2d_bytes = 0
        self.vols_per_step = args.batch * args.micro_batches * be.world
        self.w = be.torch.zeros(4)

    def step_resident(self):
        g = self.be.torch.full((4,), float(self.be.rank + 1))
        if self.be.world > 1:
            self.be.dist.all_reduce(g)                   # the "gradient exchange"
        self.w += g
        self.be._launches += 3
        return self.w.sum()

    def step_e2e(self):
        return float(self.step_resident())

    def launches_per_step(self, eager_delta_per_step):
        return int(eager_delta_per_step)

    def replicas_in_sync(self):
        if self.be.world == 1:
            return None
        both = [self.be.torch.empty_like(self.w) for _ in range(self.be.world)]
        self.be.dist.all_gather(both, self.w)
        return bool(all(self.be.torch.equal(b, both[0]) for b in both))

    def instrumented_step(self):
        t0 = time.perf_counter()
        self.step_resident()
        ms = (time.perf_counter() - t0) * 1e3
        return [("vsn_gemm_bf16", "stub", (2e9, 1e6), ms * 0.6), ("vsn_layernorm_fwd", "stub", (0, 1e6), ms * 0.4)], ms


# ------------------------------------------------------------------------------------------- native arm
def measure(be, wl, args, steps, warmup):
    """warm-up, `value` loop, host enqueue cost, `e2e` loop; every rank runs every call."""
    for _ in range(warmup):
        wl.step_resident()
    l0 = be.launch_count()
    ms = be.timed(wl.step_resident, steps)
    eager_launches = (be.launch_count() - l0) / steps
    # host-side enqueue cost of a step (no synchronisation inside): how far the step is from launch-bound
    be.sync_all()
    t0 = time.perf_counter()
    for _ in range(2):
        wl.step_resident()
    host_ms = (time.perf_counter() - t0) / 2 * 1e3
    be.sync_all()
    wl.step_e2e()
    ms_e2e = be.timed(wl.step_e2e, steps)
    return {"ms": ms, "ms_e2e": ms_e2e, "host_ms": host_ms, "launches": wl.launches_per_step(eager_launches) * steps,
            "value": wl.vols_per_step * steps / (ms / 1e3), "e2e": wl.vols_per_step * steps / (ms_e2e / 1e3)}


def native_main(args):
    be = StubBackend(args) if args.stub else CudaBackend(args)
    rank, world = be.rank, be.world
    wl = be.build(args)
    warmup = max(args.warmup, 3)
    sampler = ClockSampler(be.local)
    if rank == 0 and not args.stub:
        sampler.start()
    m = measure(be, wl, args, args.steps, warmup)
    clocks = sampler.stop() if rank == 0 and not args.stub else None
    in_sync = wl.replicas_in_sync()

    # ---- instrumented pass: per-launch CUDA-event times of every C-ABI call in one step; ALL ranks run it ---------
    roofline, table = None, []
    if not args.no_profile:
        prof, step_ms = wl.instrumented_step()
        if rank == 0:
            peaks = {}
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    peaks = json.load(f)
            except OSError:
                pass
            table, roofline, tot = summarise_profile(prof, step_ms, peaks)
            # DRAM traffic of the family's heaviest kernel from the committed ncu --set full capture
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    tr = json.load(f)
                # the headline is a kernel FAMILY over one step, so its traffic is the family's DRAM bytes per step
                # (scripts/family_traffic.py over an ncu launch list with the dram__bytes metrics), beside the
                # algorithmic bytes per step `achieved` is computed from
                ent = tr.get("family:" + roofline["kernel"])
                if ent and args.model == "swin" and not args.sam and not args.no_fuse_micro:
                    roofline["traffic"] = ent["dram_bytes_per_step"]
                    roofline["traffic_unit"] = "DRAM bytes per step, all launches of the family"
                    roofline["traffic_source"] = ent["source"]
            except OSError:
                pass
            if args.profile_out:
                os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
                with open(args.profile_out, "w") as f:
                    json.dump({"step_ms_instrumented": step_ms, "sum_of_calls_ms": tot, "calls": table}, f, indent=1)
    be.sync_all()

    # ---- extras (every rank takes part in the SAM line: it exchanges gradients) ------------------------------------
    extras = {}
    if not args.no_extras and not args.stub and args.model == "swin" and not args.sam:
        del wl
        import gc
        gc.collect()
        be.torch.cuda.empty_cache()
        a2 = argparse.Namespace(**vars(args))
        a2.sam, a2.classes, a2.mixup = True, 5, True
        w2 = be.build(a2)
        m2 = measure(be, w2, a2, max(2, args.steps // 2), 3)
        extras["swin5c_sam_ema_mixup"] = {
            "config": config_dict(a2, world), "value": round(m2["value"], 2), "unit": "volumes/s",
            "ms_per_step": round(m2["ms"] / max(2, args.steps // 2), 3),
            "e2e": {"value": round(m2["e2e"], 2), "unit": "volumes/s"}, "host_enqueue_ms_per_step": round(m2["host_ms"], 3)}
        del w2
        gc.collect()
        be.torch.cuda.empty_cache()
        be.sync_all()
        # the other half of north_star's path: ViT-3D (ViT-S, vit-3c geometry, 811 tokens, global attention on tcgen05)
        a4 = argparse.Namespace(**vars(args))
        a4.model, a4.batch, a4.classes, a4.sam, a4.mixup = "vit", 24, 3, False, False
        w4 = be.build(a4)
        m4 = measure(be, w4, a4, max(2, args.steps // 2), 3)
        extras["vit3c_adamw_ema"] = {
            "config": config_dict(a4, world), "value": round(m4["value"], 2), "unit": "volumes/s",
            "ms_per_step": round(m4["ms"] / max(2, args.steps // 2), 3),
            "e2e": {"value": round(m4["e2e"], 2), "unit": "volumes/s"},
            "model_tflops_per_gpu": round(3 * FLOP_FWD_PER_VOL["vit"] * a4.batch * a4.micro_batches
                                          / (m4["ms"] / max(2, args.steps // 2) * 1e-3) / 1e12, 1)}
        del w4
        gc.collect()
        be.torch.cuda.empty_cache()
        be.sync_all()
        if not args.no_fuse_micro and args.micro_batches > 1 and not args.torch_ddp:
            # the same step with the micro-batches accumulated one forward/backward at a time, as the reference's loop
            # does (train/train_transformer.py:1111-1190): what `micro_batches_fused` buys, stated beside the headline
            a3 = argparse.Namespace(**vars(args))
            a3.no_fuse_micro = True
            w3 = be.build(a3)
            m3 = measure(be, w3, a3, max(2, args.steps // 2), 3)
            extras["micro_batches_one_by_one"] = {
                "config": config_dict(a3, world), "value": round(m3["value"], 2), "unit": "volumes/s",
                "ms_per_step": round(m3["ms"] / max(2, args.steps // 2), 3),
                "e2e": {"value": round(m3["e2e"], 2), "unit": "volumes/s"}}
            del w3
            gc.collect()
            be.torch.cuda.empty_cache()
            be.sync_all()

        if world > 1:
            # BASELINE config 5 on N GPUs: every rank predicts its shard of the step's subjects, one all-gather at the end
            extras["swin5c_tta_ensemble_infer"] = tta_ensemble_arm(args, be.dev, be=be)
            gc.collect()
            be.torch.cuda.empty_cache()
            be.sync_all()

    # ---- baselines on rank 0 at N = 1 only -------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.stub:
        if not args.no_extras and be._lib.PRECISION == "bf16":
            # the same step in the precision mode (IEEE-half operands, GradScaler): the operand encoding is fixed per
            # process, so a child process measures it
            cmd = [sys.executable, os.path.abspath(__file__), "--precision", "f16", "--model", args.model, "--classes",
                   str(args.classes), "--steps", str(max(2, args.steps // 2)), "--warmup", "3", "--no-extras",
                   "--no-cpu-baseline", "--no-profile"] + (["--sam"] if args.sam else [])
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
                ln = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
                extras["f16_precision_mode"] = {k: ln[k] for k in ("value", "unit", "ms_per_step", "dtype", "e2e", "config")}
            except Exception as e:      # noqa: BLE001 -- an extras line must not take the headline down
                extras["f16_precision_mode"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        if not args.no_extras and args.model == "swin":
            try:
                extras["swin5c_tta_ensemble_infer"] = tta_ensemble_arm(args, be.dev)
            except RuntimeError as e:
                extras["swin5c_tta_ensemble_infer"] = {"unavailable": str(e).splitlines()[0][:200]}
            be.torch.cuda.empty_cache()
        if not args.no_extras:
            try:
                extras["eager_cuda_baseline"] = eager_cuda_arm(args)
            except (RuntimeError, ImportError) as e:
                extras["eager_cuda_baseline"] = {"unavailable": str(e).splitlines()[0][:200]}
            be.torch.cuda.empty_cache()
        if not args.no_cpu_baseline:
            cpu, _, ran = cpu_arm(args, steps=2, warmup=1)
            cpu["config_ran"] = ran

    if rank == 0:
        passes = 2 if args.sam else 1
        step_flops = 3 * FLOP_FWD_PER_VOL[args.model] * args.batch * args.micro_batches * passes
        line = {"metric": f"{args.model}3d_train_volumes_per_s", "value": round(m["value"], 2), "unit": "volumes/s",
                "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": round(m["ms"] / args.steps, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if args.stub else be._lib.PRECISION, "data": "synthetic", "config": config_dict(args, world),
                "clocks": clocks, "replicas_in_sync": in_sync, "gpu_launches": int(m["launches"]),
                "cuda_graph": bool(not args.no_graph and not args.torch_ddp),
                "host_enqueue_ms_per_step": round(m["host_ms"], 3),
                "e2e": {"value": round(m["e2e"], 2), "unit": "volumes/s", "h2d_bytes_per_step": int(wl_h2d(args)),
                        "d2h_bytes_per_step": 4, "ms_per_step": round(m["ms_e2e"] / args.steps, 3)},
                "model_tflops_per_gpu": round(step_flops / (m["ms"] / args.steps * 1e-3) / 1e12, 1),
                "roofline": roofline, "cpu_baseline": cpu, "extras": extras}
        print(json.dumps(line), flush=True)
    be.teardown()


def wl_h2d(args):
    vol = VOLUMES[args.model]
    return args.micro_batches * (args.batch * vol[0] * vol[1] * vol[2] * 2 + args.batch * args.classes * 4)


def main(argv=None):
    args = parse(argv)
    if args.precision is not None:
        os.environ["VSN_B200_PRECISION"] = args.precision     # read once, when vsn_b200._lib is imported
    if args.impl != "native":
        return reference_main(args)
    return native_main(args)


if __name__ == "__main__":
    main()
