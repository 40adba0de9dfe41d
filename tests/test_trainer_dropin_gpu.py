"""The UNMODIFIED reference trainer (train/train_transformer.py from baseline/_ref) driven through vsn_b200's drop-in
packages on a synthetic cohort: SAM(AdamW) + EMA + MixUp + balanced sampler, fp16 autocast + GradScaler, validation
with ema.apply_to / restore, asynchronous checkpoints; then the checkpoint it wrote is loaded strictly into the
reference's own model class (CPU fp32) and into the drop-in (GPU) and the logits are compared (2e-2, bf16 path).
Needs baseline/_ref (scripts/install_ref.sh; it travels with the gpurun snapshot): skipped where it is absent."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "scripts", "run_reference_trainer.py")
HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "train"))


def _run(*args, timeout=900):
    r = subprocess.run([sys.executable, HARNESS, *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout + r.stderr


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed")
@pytest.mark.parametrize("arch", ["swin", "vit"])
def test_unmodified_trainer_runs_on_the_dropin_and_checkpoints_interchange(arch, tmp_path):
    out = str(tmp_path / "run")
    log = _run("--arch", arch, "--steps", "4", "--out", out)
    assert "Using SAM optimizer" in log and "Using EMA: True" in log and "Gradient Scaler active" in log
    assert "Step 4" in log and "Checkpoint saved" in log
    log = _run("--arch", arch, "--steps", "6", "--out", out, "--resume")
    assert "Step 6" in log
    log = _run("--arch", arch, "--out", out, "--compare")
    assert "loaded strictly into the reference" in log
