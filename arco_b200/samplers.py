"""Drop-in names for the reference's helper functions, running on the GPU.

* ``grid_monte_carlo_sample`` / ``grid_as_monte_carlo_sample`` / ``monte_carlo_sample`` /
  ``as_monte_carlo_sample`` -- ``/root/reference/code/loss_helper_3d.py:35-268`` (same arguments,
  int64 result of length ``shape``), drawn by the Philox/Feistel kernel ``arco_sample_one``.  The 1-D
  variants are reached exactly like in the reference: they are what the grid samplers fall back to.
* ``dequeue_and_enqueue`` -- ``loss_helper_3d.py:12-32`` for callers that hold the adopted bank.
* ``label_onehot`` -- the trainers' override, ``train_arco_2d.py:492-498``.
"""
from __future__ import annotations

import itertools

import torch

from . import _cabi
from .bank import BankSlot

_counter = itertools.count(1)


def _device(device):
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("arco_b200 samplers run on the GPU only: no CUDA device is visible")
        device = torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _sample(func: int, high: int, shape: int, device=None, seed=None, stream_id=None) -> torch.Tensor:
    high, shape = int(high), int(shape)
    if high <= 0 or shape <= 0:
        raise ValueError("high and shape must be positive")
    dev = _device(device)
    out = torch.empty(shape, dtype=torch.int32, device=dev)
    scratch = torch.empty(shape // 2 + 4096, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        sd = int(seed) if seed is not None else int(torch.cuda.initial_seed()) & (2 ** 63 - 1)
        sid = (1 << 40) + next(_counter) if stream_id is None else int(stream_id)
        _cabi.check(_cabi.lib.arco_sample_one(func, high, shape, sd, sid, out.data_ptr(), scratch.data_ptr(),
                                              scratch.numel() * 4, torch.cuda.current_stream().cuda_stream),
                    "arco_sample_one")
    return out.long()


def grid_monte_carlo_sample(high=5233, shape=256, cut_count=4, device=None, seed=None):
    if cut_count != 4:
        raise ValueError("only cut_count=4 (the value the loss uses) is implemented")
    return _sample(_cabi.FUNC_SMC, high, shape, device, seed)


def grid_as_monte_carlo_sample(high=5233, shape=256, cut_count=4, device=None, seed=None):
    if cut_count != 4:
        raise ValueError("only cut_count=4 (the value the loss uses) is implemented")
    return _sample(_cabi.FUNC_ASMC, high, shape, device, seed)


def monte_carlo_sample(high=5233, shape=256, patch=16, device=None, seed=None):
    """1-D stratified sampler.  The loss only reaches it for ``high <= 56`` (grid fallback); for larger
    ``high`` this entry point draws with the same strata rule on the host-visible contract."""
    if patch != 16:
        raise ValueError("only patch=16 is implemented")
    if int(high) > 56:
        raise ValueError("monte_carlo_sample is only implemented as the grid sampler's fallback (high <= 56)")
    return _sample(_cabi.FUNC_SMC, high, shape, device, seed)


def as_monte_carlo_sample(high=5233, shape=256, patch=16, device=None, seed=None):
    if patch != 16:
        raise ValueError("only patch=16 is implemented")
    if int(high) > 56:
        raise ValueError("as_monte_carlo_sample is only implemented as the grid sampler's fallback (high <= 56)")
    return _sample(_cabi.FUNC_ASMC, high, shape, device, seed)


@torch.no_grad()
def dequeue_and_enqueue(keys, queue, queue_ptr, queue_size):
    """FIFO append of ``keys`` to ``queue`` (``memobank[c]``), keeping the newest ``queue_size`` rows.
    Works on an adopted :class:`BankSlot` (rows stay in HBM) and on a plain ``[tensor]`` list."""
    n = int(keys.shape[0])
    if isinstance(queue, BankSlot):
        bank, c = queue.bank, queue.cls
        merged = torch.cat((bank.rows_of(c), keys.detach().to(bank.device, torch.float32)), dim=0)
        full = merged.shape[0] >= queue_size
        bank.replace_rows(c, merged)
        ptr = queue_size if full else (int(queue_ptr) + n) % queue_size
        bank.ptr[c] = ptr
        bank._ptr_alias[c] = ptr
    else:
        merged = torch.cat((queue[0], keys.detach().to(queue[0].device)), dim=0)
        full = merged.shape[0] >= queue_size
        queue[0] = merged[-queue_size:, :] if full else merged
        ptr = queue_size if full else (int(queue_ptr) + n) % queue_size
    queue_ptr[0] = ptr
    return n


def label_onehot(inputs, num_segments):
    """int labels ``[B,*S]`` -> float32 one-hot ``[B,C,*S]`` on the input's device, -1 -> class 0."""
    if not inputs.is_cuda:
        raise RuntimeError("arco_b200.label_onehot needs a CUDA tensor: there is no CPU fallback")
    lab = inputs.to(torch.int64).contiguous()
    B = lab.shape[0]
    S = lab[0].numel()
    out = torch.empty((B, int(num_segments)) + tuple(lab.shape[1:]), dtype=torch.float32, device=lab.device)
    with torch.cuda.device(lab.device):
        _cabi.check(_cabi.lib.arco_label_onehot(lab.data_ptr(), out.data_ptr(), B, int(num_segments), S,
                                                torch.cuda.current_stream().cuda_stream), "arco_label_onehot")
    return out
