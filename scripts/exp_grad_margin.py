#!/usr/bin/env python
"""GPU experiment: worst parameter-gradient error of the parity cases with the LayerNorm-backward inputs in bf16 vs fp32
(swin.LN_DY_DTYPE), and the step time of each setting.  Prints one line per (setting, case)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import vsn_b200  # noqa: F401,E402
from vsn_b200 import swin  # noqa: E402
from tests import test_model_gpu as T  # noqa: E402

for dt in (torch.bfloat16, torch.float32):
    swin.LN_DY_DTYPE = dt
    for fn, args in ((T.test_swin_matches_reference, ("swin_tiny_odd",)), (T.test_swin_matches_reference, ("swin_small_even",)),
                     (T.test_swin5c_full_size_train_gradients, ()), (T.test_vit3c_full_size_train_gradients, ())):
        try:
            fn(*args)
            print("OK  ", dt, fn.__name__, args, flush=True)
        except AssertionError as e:
            print("FAIL", dt, fn.__name__, args, str(e)[:300], flush=True)
