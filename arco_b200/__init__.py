"""arco_b200 -- B200-native (sm_100a) implementation of ARCO's stratified contrastive loss.

Drop-in for the one hot path of charlesyou999648/ARCO: ``compute_contra_memobank_loss`` and its
samplers / memory-bank enqueue (``code/loss_helper_3d.py:12-513``, ``code/loss_helper.py:142-686``).
A trainer switches with one import line::

    from arco_b200.loss_helper_3d import *     # train_arco_2d.py:24  (4-D image tensors)
    from arco_b200.loss_helper import *        # train_arco_3d.py:22  (5-D volume tensors)

The public names below are resolved on first use and load ``arco_b200/lib/libarco_b200.so`` through ctypes;
if the library is missing or stale that raises (there is no CPU / PyTorch fallback).  Only
``arco_b200.build`` -- the nvcc recipe that produces the library -- is importable without it.
"""
import importlib

_EXPORTS = {
    "ArcoError": "_cabi", "version": "_cabi",
    "BankSlot": "bank", "DeviceMemoryBank": "bank", "synchronize_bank": "bank",
    "compute_contra_memobank_loss": "contra", "compute_contra_memobank_loss_from_logits": "contra",
    "compute_contra_memobank_loss_from_features": "producers", "FeatureExtractor": "producers", "FeatureExtractor_3d": "producers", "make_q_representation": "producers",
    "prepare_contrast_inputs": "prepare", "softmax_entropy": "prepare", "entropy_masks": "prepare",
    "dense_similarity": "similarity",
    "compute_unsupervised_loss": "stepterms", "RandTPS": "stepterms", "tps_equivariance_loss": "stepterms",
    "get_revisiting_loss": "revisit", "revisit_enqueue": "revisit", "_dequeue_and_enqueue": "revisit",
    "as_monte_carlo_sample": "samplers", "dequeue_and_enqueue": "samplers", "grid_as_monte_carlo_sample": "samplers",
    "grid_monte_carlo_sample": "samplers", "label_onehot": "samplers", "monte_carlo_sample": "samplers",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module 'arco_b200' has no attribute {name!r}")
    value = getattr(importlib.import_module(f"{__name__}.{mod}"), name)
    globals()[name] = value
    return value


def __dir__():
    return sorted(list(globals()) + list(_EXPORTS))
