"""Drop-in for utils/ema.py:10-178."""
from vsn_b200.optim import EMAModel  # noqa: F401
