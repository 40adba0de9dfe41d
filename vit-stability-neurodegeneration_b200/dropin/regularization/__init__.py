"""`regularization` package of the reference with SAM replaced by the multi-tensor CUDA implementation."""
import vsn_b200  # noqa: F401
from vsn_b200.dropin._chain import chain as _chain

_chain(globals(), "regularization")
from .sam import SAM  # noqa: E402,F401  (also when no reference package is on the path)
