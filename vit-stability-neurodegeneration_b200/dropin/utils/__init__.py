"""`utils` package of the reference with EMAModel replaced by the on-device snapshot ring."""
import vsn_b200  # noqa: F401
from vsn_b200.dropin._chain import chain as _chain

_chain(globals(), "utils")
from .ema import EMAModel  # noqa: E402,F401
