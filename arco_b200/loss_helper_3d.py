"""Star-importable twin of the reference's ``code/loss_helper_3d.py`` -- the **2-D image** variant
(rep ``[B,D,H,W]``; the reference's file names are inverted) that
``train_arco_2d.py:24`` imports with ``from loss_helper_3d import *``."""
from .contra import compute_contra_memobank_loss  # noqa: F401
from .samplers import (  # noqa: F401
    as_monte_carlo_sample,
    dequeue_and_enqueue,
    grid_as_monte_carlo_sample,
    grid_monte_carlo_sample,
    label_onehot,
    monte_carlo_sample,
)

__all__ = [
    "as_monte_carlo_sample", "compute_contra_memobank_loss", "dequeue_and_enqueue",
    "grid_as_monte_carlo_sample", "grid_monte_carlo_sample", "label_onehot", "monte_carlo_sample",
]
