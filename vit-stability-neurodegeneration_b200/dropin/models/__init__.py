"""`models` package of the reference with the Swin-3D and ViT-3D modules replaced by the vsn_b200 path."""
import vsn_b200  # noqa: F401  (registers the package; fails loudly if the repo root is not importable)
from vsn_b200.dropin._chain import chain as _chain

_chain(globals(), "models")
