"""arco_b200 -- B200-native (sm_100a) implementation of ARCO's stratified contrastive loss.

Drop-in for the one hot path of charlesyou999648/ARCO: ``compute_contra_memobank_loss`` and its
samplers / memory-bank enqueue (``code/loss_helper_3d.py:12-513``, ``code/loss_helper.py:142-686``).
A trainer switches with one import line::

    from arco_b200.loss_helper_3d import *     # train_arco_2d.py:24  (4-D image tensors)
    from arco_b200.loss_helper import *        # train_arco_3d.py:22  (5-D volume tensors)

Importing the package loads ``arco_b200/lib/libarco_b200.so``; there is no CPU fallback.
"""
from . import _cabi
from ._cabi import ArcoError, version
from .bank import BankSlot, DeviceMemoryBank, synchronize_bank
from .contra import compute_contra_memobank_loss
from .samplers import (
    as_monte_carlo_sample,
    dequeue_and_enqueue,
    grid_as_monte_carlo_sample,
    grid_monte_carlo_sample,
    label_onehot,
    monte_carlo_sample,
)

__all__ = [
    "ArcoError", "BankSlot", "DeviceMemoryBank", "as_monte_carlo_sample", "compute_contra_memobank_loss",
    "dequeue_and_enqueue", "grid_as_monte_carlo_sample", "grid_monte_carlo_sample", "label_onehot",
    "monte_carlo_sample", "synchronize_bank", "version",
]
