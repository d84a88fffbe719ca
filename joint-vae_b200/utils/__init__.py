"""Counterparts of the reference's utils/ that sit next to the hot path."""
from . import batch_loader, roc_curves, save_load      # noqa: F401
