"""The shadow packages resolve the reference's import paths to the vsn_b200 classes and leave every other module
of the reference reachable.  Needs the reference checkout, so it only runs where /root/reference exists."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VSN_REFERENCE_ROOT", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not present")
def test_reference_import_paths_resolve_to_vsn_b200():
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        from oracle import refshim
        root = refshim.install()                      # timm / monai stand-ins (absent from this image)
        sys.path.remove(root)
        sys.path.insert(0, {os.path.join(ROOT, 'vit-stability-neurodegeneration_b200', 'dropin')!r})
        sys.path.append(root)                         # the trainer appends its root last
        from models.swin_transformer_3d import SwinTransformerT
        from models.vit_3d import ViTS
        from regularization import LabelSmoothingLoss, SAM
        from utils import EMAModel, get_params_groups, cosine_scheduler
        import vsn_b200.swin_model, vsn_b200.vit_model, vsn_b200.optim
        assert SwinTransformerT is vsn_b200.swin_model.SwinTransformerT
        assert ViTS is vsn_b200.vit_model.ViTS
        assert SAM is vsn_b200.optim.SAM and EMAModel is vsn_b200.optim.EMAModel
        assert LabelSmoothingLoss.__module__ == 'regularization.label_smoothing'
        assert get_params_groups.__module__ == 'utils.helper'
        print('ok')
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
