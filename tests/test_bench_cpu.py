"""bench.py's control flow under torchrun with 2 ranks on CPU (gloo): the stub step contains a collective, as the real
step contains the gradient all-reduce, so any call that only some ranks make (round 1: the instrumented pass ran on
rank 0 alone and deadlocked every N > 1 run) hangs this test instead of the driver's scaling run."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, *bench_args, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"),
           "--gpus", str(nproc), *bench_args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout                      # rank 0 alone prints, exactly one JSON line
    return json.loads(lines[0])


@pytest.mark.timeout(300)
def test_bench_control_flow_two_ranks_gloo():
    line = _torchrun(2, "--stub", "--steps", "3", "--warmup", "1")
    assert line["n_gpus"] == 2 and line["steps"] == 3 and line["warmup"] >= 3
    assert line["replicas_in_sync"] is True
    assert line["scaling"] == "weak" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] > 0
    rf = line["roofline"]                                  # the instrumented pass ran (on both ranks) and was summarised
    assert rf["kernel"] == "gemm" and set(rf["families"]) == {"gemm", "layernorm"}
    assert 0 < rf["frac"] and rf["bound"] in ("tensor", "hbm")
    assert line["cpu_baseline"] is None                    # N > 1: no CPU leg


@pytest.mark.timeout(300)
def test_bench_control_flow_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--stub", "--steps", "2"], capture_output=True,
                       text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert line["n_gpus"] == 1 and line["replicas_in_sync"] is None and line["gpu_launches"] == 6


@pytest.mark.timeout(600)
def test_reference_arm_under_torchrun_runs_on_rank0_only():
    """`--impl reference` launched like the native arm: rank 0 times the reference's CPU path (the unmodified
    reference where baseline/_ref or the checkout exists, else the oracle port), the other rank exits 0."""
    line = _torchrun(2, "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-batch", "1", timeout=580)
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "volumes/s"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["config_ran"]["micro_batch"] == 1 and line["config_ran"]["parallelism"] == "cpu"
    assert line["e2e"] == {"value": line["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
