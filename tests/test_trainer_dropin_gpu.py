"""The UNMODIFIED reference trainer (train/train_transformer.py from baseline/_ref) driven through vsn_b200's drop-in
packages on a synthetic cohort: SAM(AdamW) + EMA + MixUp + balanced sampler, fp16 autocast + GradScaler, validation
with ema.apply_to / restore, asynchronous checkpoints; then the checkpoint it wrote is loaded strictly into the
reference's own model class (CPU fp32) and into the drop-in (GPU) and the logits are compared (2e-2, bf16 path);
then the UNMODIFIED evaluation script (eval/eval_transformer.py) runs over those checkpoints through the drop-in and its
prediction CSVs are compared with a control run on the reference's own modules.
Needs baseline/_ref (scripts/install_ref.sh; it travels with the gpurun snapshot): skipped where it is absent."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "scripts", "run_reference_trainer.py")
HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "train"))


def _run(*args, timeout=900, env=None):
    r = subprocess.run([sys.executable, HARNESS, *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                       env=None if env is None else dict(os.environ, **env))
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
    # the UNMODIFIED evaluation script (eval/eval_transformer.py) over the checkpoints that run wrote: through the drop-in
    # (fp32 call and fp16-autocast call), then the control run with the reference's own modules, then the prediction
    # CSVs of the two against each other (class probabilities, bf16 path: 2e-2 absolute)
    fast = ("--eval-last-only", "--bootstrap", "20")
    log = _run("--arch", arch, "--out", out, "--eval", *fast)
    assert "Saved in-domain predictions" in log and "Test (ID)" in log
    if arch == "swin":
        log = _run("--arch", arch, "--out", out, "--eval", "--use-amp", "--eval-folder", "eval_dropin_amp", *fast)
        assert "Saved in-domain predictions" in log
    log = _run("--arch", arch, "--out", out, "--eval", "--reference-models", *fast)
    assert "Saved in-domain predictions" in log
    log = _run("--arch", arch, "--out", out, "--eval-compare")
    assert "eval-compare: ok" in log


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed")
def test_unmodified_trainer_runs_in_the_f16_precision_mode(tmp_path):
    """VSN_B200_PRECISION=f16: the trainer's own fp16 loop (autocast + GradScaler, SAM's unscale_ / second_step(scaler))
    drives the IEEE-half build of the kernels; the checkpoint still interchanges with the reference model."""
    out = str(tmp_path / "run16")
    env = {"VSN_B200_PRECISION": "f16"}
    log = _run("--arch", "swin", "--steps", "4", "--out", out, env=env)
    assert "Using SAM optimizer" in log and "Gradient Scaler active" in log and "Step 4" in log and "Checkpoint saved" in log
    import re
    vals = [float(v) for v in re.findall(r"Val: loss ([0-9.naninf]+),", log)]
    assert vals and all(v == v and v < 10.0 for v in vals), vals          # finite validation losses
    log = _run("--arch", "swin", "--out", out, "--compare", env=env)
    assert "loaded strictly into the reference" in log
