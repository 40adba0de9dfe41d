"""Drop-in for regularization/sam.py:9-165."""
from vsn_b200.optim import SAM  # noqa: F401
