"""Nearest-neighbour "revisiting" loss and its random-pool queue.

Drop-ins for the reference trainer's helpers (``/root/reference/code/train_arco_2d.py``):

* :func:`get_revisiting_loss` -- ``get_revisiting_loss(random_pool, rep_u, rep_u_teacher, topk=5)`` (:126-136)
* :func:`_dequeue_and_enqueue` -- ``_dequeue_and_enqueue(keys, queue, queue_ptr)`` (:109-120), and the fused
  :func:`revisit_enqueue` that replaces the trainer's three lines :400-402 (view, normalize, enqueue).

The reference normalises both ``[bs, D*H*W]`` tensors (two read+write passes each), runs two fp32 einsums against the
4.7 GB pool and a third normalise pass for the enqueue.  Here one streaming CUDA pass reads ``rep_u``,
``rep_u_teacher`` and ``random_pool`` exactly once (``arco_revisit_loss``), and the enqueue is one scale-and-copy pass
that reuses the norms of the loss pass (``arco_revisit_enqueue``).  No CPU path.

As in the reference the loss carries no gradient to the optimised parameters: the neighbours come from the student
(``topk`` indices are not differentiable) and the averaged distances from the teacher, whose extractor is frozen
(train_arco_2d.py:250-253).  ``get_revisiting_loss`` therefore returns a constant (detached) scalar.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi

_LAST = {}          # id(random_pool storage) -> stats of the last loss pass (norms for the fused enqueue)


def _flat2(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    return t.view(t.shape[0], -1)


def get_revisiting_loss(random_pool: torch.Tensor, rep_u: torch.Tensor, rep_u_teacher: torch.Tensor, topk: int = 5,
                        return_index: bool = False):
    """Same arguments and value as the reference (train_arco_2d.py:126-136).  ``random_pool``: f32 ``[K, D*H*W]`` unit
    rows on the GPU; ``rep_u`` / ``rep_u_teacher``: ``[bs, D, H, W]`` (or already ``[bs, L]``) f32 or bf16."""
    if not (rep_u.is_cuda and rep_u_teacher.is_cuda and random_pool.is_cuda):
        raise RuntimeError("arco_b200.get_revisiting_loss needs CUDA tensors: there is no CPU fallback")
    if rep_u.dtype not in (torch.float32, torch.bfloat16) or rep_u_teacher.dtype != rep_u.dtype:
        raise ValueError("rep_u / rep_u_teacher must both be float32 or both bfloat16")
    if random_pool.dtype != torch.float32 or random_pool.dim() != 2 or not random_pool.is_contiguous():
        raise ValueError("random_pool must be a contiguous float32 [K, D*H*W] tensor")
    s, t = _flat2(rep_u), _flat2(rep_u_teacher)
    bs, L = s.shape
    K = random_pool.shape[0]
    if t.shape != s.shape or random_pool.shape[1] != L:
        raise ValueError(f"shapes disagree: rep_u {tuple(s.shape)}, rep_u_teacher {tuple(t.shape)}, random_pool {tuple(random_pool.shape)}")
    if not 1 <= topk <= K:
        raise ValueError("topk must be in [1, K]")
    dev = s.device
    with torch.cuda.device(dev):
        out = torch.empty(1 + 2 * bs * K + 2 * bs, dtype=torch.float32, device=dev)
        nn_index = torch.empty((bs, topk), dtype=torch.int32, device=dev)
        scratch = torch.empty(int(_cabi.lib.arco_revisit_scratch_bytes(bs, K)), dtype=torch.uint8, device=dev)
        stats = out[1:]
        _cabi.check(_cabi.lib.arco_revisit_loss(
            s.data_ptr(), t.data_ptr(), random_pool.data_ptr(), bs, K, L,
            _cabi.BF16 if s.dtype == torch.bfloat16 else _cabi.F32, int(topk), out.data_ptr(), nn_index.data_ptr(),
            stats.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "arco_revisit_loss")
    _LAST[random_pool.data_ptr()] = (stats, rep_u_teacher.data_ptr(), bs, K, L)
    loss = out[0]
    return (loss, nn_index.long()) if return_index else loss


def revisit_enqueue(rep_u_teacher: torch.Tensor, random_pool: torch.Tensor, random_pool_ptr: torch.Tensor) -> None:
    """The trainer's lines :400-402 in one pass: ``random_pool[ptr:ptr+bs] = normalize(rep_u_teacher.view(bs, -1))`` and the
    pointer update of ``_dequeue_and_enqueue`` (:109-120).  Reuses the norms of the preceding ``get_revisiting_loss`` call
    on the same tensors when there was one."""
    t = _flat2(rep_u_teacher)
    bs, L = t.shape
    K = random_pool.shape[0]
    if K % bs != 0:
        raise AssertionError("args.K % batch_size == 0")                     # train_arco_2d.py:113
    dev = t.device
    last = _LAST.get(random_pool.data_ptr())
    with torch.cuda.device(dev):
        if last is not None and last[1:] == (rep_u_teacher.data_ptr(), bs, K, L):
            stats = last[0]
        else:
            stats = torch.zeros(2 * bs * K + 2 * bs, dtype=torch.float32, device=dev)
            stats[2 * bs * K + bs:] = t.float().pow(2).sum(dim=1)
        ptr = int(random_pool_ptr.reshape(-1)[0])
        _cabi.check(_cabi.lib.arco_revisit_enqueue(
            t.data_ptr(), stats.data_ptr(), random_pool.data_ptr(), ptr, bs, K, L,
            _cabi.BF16 if t.dtype == torch.bfloat16 else _cabi.F32, torch.cuda.current_stream().cuda_stream),
            "arco_revisit_enqueue")
    random_pool_ptr[0] = (ptr + bs) % K


@torch.no_grad()
def _dequeue_and_enqueue(keys: torch.Tensor, queue: torch.Tensor, queue_ptr: torch.Tensor, K: Optional[int] = None) -> None:
    """Reference signature (train_arco_2d.py:109-120): ``keys`` are already normalised ``[bs, L]`` rows."""
    bs = keys.shape[0]
    K = int(K if K is not None else queue.shape[0])
    ptr = int(queue_ptr.reshape(-1)[0])
    assert K % bs == 0
    queue[ptr: ptr + bs] = keys.to(queue.dtype)
    queue_ptr[0] = (ptr + bs) % K
