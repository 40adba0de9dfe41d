#!/usr/bin/env python
"""Headline benchmark: Swin-3D training volumes/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the oracle port on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # N > 1, one rank per GPU over NCCL

A "step" is one optimiser step of the reference's hot loop (train/train_transformer.py:1104-1298) on synthetic
MNI-shaped fp16 volumes: `--micro-batches` accumulation micro-batches of `--batch` volumes per GPU (default
2 x 8 = the reference's per-GPU schedule at 8 GPUs, EFFECTIVE_BATCH_SIZE 128), soft-target CE, AdamW(fused) and
the EMA update; `--sam` adds the SAM two-pass step (BASELINE config 3).  Scaling is weak (per-GPU work fixed).

  value      whole-job volumes/s with the inputs already resident in HBM
  e2e        same loop through the public module API with pinned HOST inputs: H2D copies of every micro-batch and
             the D2H read of the loss are inside the timed region
  roofline   the kernel group with the largest share of the step (per-launch CUDA-event times, measured in an
             instrumented pass of the same step) against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/swin3d_oracle.py, a port of the reference's path) timed on the host cores
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
VOLUME = (144, 168, 144)
FLOP_FWD_PER_VOL = 108.73e9          # Swin-T forward on the padded grids the reference runs


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="volumes per micro-batch per GPU (BATCH_SIZE)")
    ap.add_argument("--micro-batches", type=int, default=2, help="accumulation micro-batches per step per GPU")
    ap.add_argument("--classes", type=int, default=3)
    ap.add_argument("--sam", action="store_true", help="SAM two-pass step (BASELINE config 3)")
    ap.add_argument("--no-ema", action="store_true")
    ap.add_argument("--torch-ddp", action="store_true", help="N>1: use torch DDP instead of vsn_b200.ddp.GradAllReduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the instrumented pass (roofline = null)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel table of the instrumented pass here")
    ap.add_argument("--cpu-batch", type=int, default=2)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- synthetic data
def synth_batch(batch, classes, seed, mixup=True):
    """fp16 [B,1,D,H,W] MNI-like volumes + soft labels, generated on the host."""
    import torch
    g = torch.Generator().manual_seed(seed)
    D, H, W = VOLUME
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


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_arm(args, steps, warmup):
    """The oracle port of the reference path on the host cores: fwd + bwd + AdamW + EMA at batch `--cpu-batch`."""
    import numpy as np
    import torch
    from oracle import swin3d_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    shapes = swin_state_shapes(args.classes)
    sd = {}
    for k, s in shapes.items():
        leaf = k.rsplit(".", 1)[-1]
        if "norm" in k and leaf == "weight":
            t = torch.ones(s)
        elif leaf == "bias":
            t = torch.zeros(s)
        else:
            t = torch.nn.init.trunc_normal_(torch.empty(s), std=0.02)
        sd[k] = t.requires_grad_(True)
    opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, weight_decay=0.05)
    snaps = [[v.detach().numpy().copy() for v in sd.values()]]
    x16, y = synth_batch(args.cpu_batch, args.classes, seed=99)
    x = x16.float()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        z = O.swin_forward(sd, x, patch=SWIN["patch_size"], window=SWIN["window_size"], depths=SWIN["depths"],
                           heads=SWIN["num_heads"])
        O.soft_target_ce(z, y, 0.1).backward()
        opt.step()
        snaps.append([v.detach().numpy().copy() for v in sd.values()])
        snaps = snaps[-3:]
        _ = [O.ema_average([s[i] for s in snaps], 0.999) for i in range(len(snaps[0]))]
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": args.cpu_batch / sec, "unit": "volumes/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} optimiser steps of {args.cpu_batch} volumes {VOLUME} (fwd+bwd+AdamW+EMA), fp32, "
                      f"oracle/swin3d_oracle.py, {sec:.2f} s/step"}, sec


def swin_state_shapes(classes):
    """state_dict shapes of Swin-T at the benchmark geometry (parameter container built on the host)."""
    import torch
    import vsn_b200  # noqa: F401
    from vsn_b200.swin_model import SwinTransformerT
    m = SwinTransformerT(in_channels=1, mlp_ratio=4.0, qkv_bias=True, dropout=0.0, attention_dropout=0.0,
                         stochastic_depth_prob=0.0, num_classes=classes, norm_layer=torch.nn.LayerNorm, **SWIN)
    return {k: tuple(v.shape) for k, v in m.state_dict().items() if v.is_floating_point()}


def config_dict(args, world):
    return {"workload": f"Swin-3D (Swin-T, swin-{args.classes}c geometry) "
                        f"{'SAM(AdamW)' if args.sam else 'AdamW'}{'' if args.no_ema else '+EMA'} training step",
            "volume": list(VOLUME), "micro_batch": args.batch, "micro_batches_per_step": args.micro_batches,
            "global_batch": args.batch * args.micro_batches * world, "parallelism": f"dp{world}",
            "l2": "per-step inputs (111 MB) and activations (GBs) exceed the 126 MB L2; no explicit flush"}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, sec = cpu_arm(args, max(1, args.steps), max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": "swin3d_train_volumes_per_s", "value": cb["value"], "unit": "volumes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, 1), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return reference_main(args)

    import torch
    import torch.distributed as dist
    import vsn_b200  # noqa: F401
    from vsn_b200 import _lib
    from vsn_b200.swin_model import SwinTransformerT
    from vsn_b200.train import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the vsn_b200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    model = SwinTransformerT(in_channels=1, mlp_ratio=4.0, qkv_bias=True, dropout=0.0, attention_dropout=0.0,
                             stochastic_depth_prob=0.15, num_classes=args.classes, norm_layer=torch.nn.LayerNorm,
                             **SWIN).to(dev)
    model.train()
    ddp, sync = None, None
    if world > 1:
        if args.torch_ddp:
            ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], output_device=local,
                                                            gradient_as_bucket_view=True)
        else:
            from vsn_b200.ddp import GradAllReduce
            sync = GradAllReduce(model.parameters(), bucket_mb=25.0)
    ts = TrainStep(model, use_sam=args.sam, use_ema=not args.no_ema, ddp_model=ddp, grad_sync=sync)

    G = args.micro_batches
    host = [synth_batch(args.batch, args.classes, seed=1234 + rank * 100 + i) for i in range(G)]
    host = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    resident = [(x.to(dev), y.to(dev)) for x, y in host]
    h2d_bytes = sum(x.numel() * x.element_size() + y.numel() * y.element_size() for x, y in host)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms

    # ---- resident-input loop (value) -----------------------------------------------------------
    def step_resident():
        ts.step(resident)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - l0
    # host-side enqueue cost of a step (no synchronisation inside): tells how far the step is from launch-bound
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        step_resident()
    host_ms = (time.perf_counter() - t0) / 2 * 1e3
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    vols_per_step = args.batch * G * world
    value = vols_per_step * args.steps / (ms / 1e3)

    # ---- end-to-end loop: pinned host inputs, H2D per micro-batch, D2H of the loss -----------------
    # Two sets of staging buffers: the H2D copies of step i+1 are enqueued on a copy stream right after step i's
    # kernels, so they overlap its compute; every timed step still contains exactly one H2D of all its inputs
    # (the copy for the first timed step is issued by the preceding warm-up step, the last timed step issues one
    # for a step that is never run) and the D2H read of its loss.
    copy_stream = torch.cuda.Stream()
    bufs = [[[torch.empty_like(x, device=dev), torch.empty_like(y, device=dev)] for x, y in host] for _ in range(2)]
    evs = [[torch.cuda.Event() for _ in host] for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]      # compute that read staging set s has been enqueued/finished
    state = {"cur": 0, "primed": False}

    def enqueue_h2d(slot):
        copy_stream.wait_event(done[slot])              # the step that last read this staging set is finished
        with torch.cuda.stream(copy_stream):
            for (hx, hy), (dx, dy), ev in zip(host, bufs[slot], evs[slot]):
                dx.copy_(hx, non_blocking=True)
                dy.copy_(hy, non_blocking=True)
                ev.record(copy_stream)

    def step_e2e():
        main_stream = torch.cuda.current_stream()
        cur = state["cur"]
        if not state["primed"]:
            done[0].record(main_stream)
            done[1].record(main_stream)
            enqueue_h2d(cur)
            state["primed"] = True
        batches = []
        for (dx, dy), ev in zip(bufs[cur], evs[cur]):
            main_stream.wait_event(ev)
            batches.append((dx, dy))
        loss = ts.step(batches)
        done[cur].record(main_stream)
        enqueue_h2d(cur ^ 1)                            # next step's inputs travel while this step computes
        state["cur"] = cur ^ 1
        return float(loss.item())                     # D2H read of the step's result

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = vols_per_step * args.steps / (ms_e2e / 1e3)

    # N > 1: every rank saw different volumes, so identical weights after all these steps mean the gradient exchange
    # delivered the same mean gradient everywhere (checked bit for bit on a fp64 checksum of all parameters)
    in_sync = None
    if world > 1:
        cks = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        allc = [torch.empty_like(cks) for _ in range(world)]
        dist.all_gather(allc, cks)
        in_sync = bool(all(torch.equal(c, allc[0]) for c in allc))
        if not in_sync:
            raise RuntimeError(f"replicas diverged: parameter checksums {[float(c) for c in allc]}")

    # ---- instrumented pass: per-launch CUDA-event times of every C-ABI call in one step -----------
    roofline, table = None, []
    if not args.no_profile and rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "measured (MEASURED_PEAKS.json, sustained bf16 / copy bandwidth)" if peaks else "fallback (B200_PROFILING.md)"
        _lib.PROFILE = []
        t_e0, t_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_e0.record()
        step_resident()
        t_e1.record()
        torch.cuda.synchronize()
        prof, _lib.PROFILE = _lib.PROFILE, None
        step_ms = t_e0.elapsed_time(t_e1)
        groups = {}
        for name, tag, work, s, e in prof:
            g = groups.setdefault((name, tag), {"n": 0, "ms": 0.0, "flops": work[0], "bytes": work[1]})
            g["n"] += 1
            g["ms"] += s.elapsed_time(e)
        tot = sum(g["ms"] for g in groups.values())
        for (name, tag), g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
            per = g["ms"] / g["n"]
            row = {"call": name, "shape": tag, "launches": g["n"], "ms_total": round(g["ms"], 4),
                   "ms_per_call": round(per, 5), "share": round(g["ms"] / tot, 4)}
            if g["flops"]:
                row["tflops"] = round(g["flops"] / (per * 1e-3) / 1e12, 2)
            if g["bytes"]:
                row["gbs"] = round(g["bytes"] / (per * 1e-3) / 1e9, 1)
            table.append(row)
        by_call = {}
        for r in table:
            by_call[r["call"]] = by_call.get(r["call"], 0.0) + r["ms_total"]
        top = table[0]
        (name, tag) = next(k for k, g in groups.items() if k[0] == top["call"] and k[1] == top["shape"])
        g = groups[(name, tag)]
        per_s = g["ms"] / g["n"] * 1e-3
        if g["flops"]:
            ach = g["flops"] / per_s / 1e12
            roofline = {"bound": "tensor", "achieved": round(ach, 2), "peak": tc_peak, "unit": "TFLOP/s",
                        "frac": round(ach / tc_peak, 4), "traffic": None}
        else:
            ach = g["bytes"] / per_s / 1e9
            roofline = {"bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(ach / hbm_peak, 4), "traffic": None}
        roofline.update({"kernel": name, "shape": tag, "launches_per_step": g["n"],
                         "ms_per_launch": round(per_s * 1e3, 5), "share_of_step": round(g["ms"] / step_ms, 4),
                         "peak_source": peak_src,
                         "share_by_call": {k: round(v / tot, 4) for k, v in sorted(by_call.items(), key=lambda kv: -kv[1])},
                         "instrumented_step_ms": round(step_ms, 3)})
        # the metric's second half: window-attention TFLOP/s vs peak (stage-0 kernels, fwd and bwd)
        wa = {}
        for r in table:
            if r["call"] in ("vsn_attn_fwd", "vsn_attn_bwd") and "tflops" in r:
                key = r["call"].replace("vsn_attn_", "") + (":shifted" if "mask=1" in r["shape"] else ":plain")
                if key not in wa or r["ms_total"] > wa[key]["ms_total"]:
                    wa[key] = {"shape": r["shape"], "tflops": r["tflops"], "frac_of_peak": round(r["tflops"] / tc_peak, 4),
                               "ms_total": r["ms_total"]}
        roofline["window_attention"] = wa
        # HBM-bound kernel families of the path (LayerNorm, casts, SAM/EMA): achieved GB/s of the heaviest shape
        hb = {}
        for r in table:
            if r["call"] in ("vsn_layernorm_fwd", "vsn_layernorm_bwd", "vsn_cast_rows_bf16", "vsn_mt_ema", "vsn_mt_cast_bf16",
                             "vsn_merge_gather", "vsn_patch_gather") and "gbs" in r:
                if r["call"] not in hb or r["ms_total"] > hb[r["call"]]["ms_total"]:
                    hb[r["call"]] = {"shape": r["shape"], "gbs": r["gbs"], "frac_of_hbm_peak": round(r["gbs"] / hbm_peak, 3),
                                     "ms_total": r["ms_total"]}
        roofline["hbm_kernels"] = hb
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/ncu_traffic.json)
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)
            ent = tr.get(f"{name}|{tag}")
            if ent:
                roofline["traffic"] = ent["dram_bytes_per_launch"]
                roofline["traffic_source"] = ent["source"]
        except OSError:
            pass
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            with open(args.profile_out, "w") as f:
                json.dump({"step_ms_instrumented": step_ms, "sum_of_calls_ms": tot, "calls": table}, f, indent=1)

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_arm(args, steps=2, warmup=1)

    if rank == 0:
        step_flops = 3 * FLOP_FWD_PER_VOL * args.batch * G * (2 if args.sam else 1)
        line = {"metric": "swin3d_train_volumes_per_s", "value": round(value, 2), "unit": "volumes/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_dict(args, world),
                "clocks": clocks, "replicas_in_sync": in_sync, "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_ms, 3),
                "e2e": {"value": round(e2e_value, 2), "unit": "volumes/s", "h2d_bytes_per_step": int(h2d_bytes),
                        "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
                "model_tflops_per_gpu": round(step_flops / (ms / args.steps * 1e-3) / 1e12, 1),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
