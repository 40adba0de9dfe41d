"""Host side of the input step (vsn_b200/data.py) against tests/golden/mixup.npz, written by oracle/make_golden_input.py
from the UNMODIFIED reference's MRIMixUp (dataset/dataset.py:186-286, seeded branch): who is mixed with whom is index
work -- bit-exact -- and the soft labels follow from the Beta draw."""
import importlib.util
import os
import sys

import numpy as np

from tests.helpers import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plan_fn():
    # data.py imports the CUDA ops lazily through the package; mixup_plan itself is pure numpy
    import vsn_b200  # noqa: F401
    from vsn_b200 import data
    return data.mixup_plan


def test_mixup_plan_reproduces_reference_partners_and_labels():
    g = golden("mixup")
    names = ["AD", "CN", "FTD"]
    diagnoses = [names[int(i)] for i in g["diagnoses"]]
    class_list = sorted(set(diagnoses))
    class_indices = {c: [i for i, d in enumerate(diagnoses) if d == c] for c in class_list}
    plan = _plan_fn()
    for epoch in (0, 3):
        mixed = 0
        for idx in range(len(diagnoses)):
            partner, a = plan(idx, seed=7, epoch=epoch, diagnoses=diagnoses, class_indices=class_indices,
                              class_list=class_list, alpha=0.3, mixup_prob=0.7)
            assert (partner if partner is not None else -1) == int(g[f"partner_e{epoch}"][idx])      # bit-exact
            y = g["labels"][idx] if partner is None else \
                np.float32(a) * g["labels"][idx] + np.float32(1.0 - a) * g["labels"][partner]
            np.testing.assert_allclose(y, g[f"y_e{epoch}"][idx], rtol=1e-6, atol=1e-7)
            mixed += partner is not None
            if partner is not None:
                assert diagnoses[partner] != diagnoses[idx]        # always a DIFFERENT class (dataset.py:252-257)
        assert 0 < mixed < len(diagnoses)
