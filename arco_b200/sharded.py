"""Batch-sharded (multi-GPU) use of the loss -- host-side helpers.

The reference has no multi-GPU loss (``nn.DataParallel`` gathers everything on device 0,
``/root/reference/code/model_2D.py:188-198``).  The path shards naturally over the batch (SURVEY.md
section 8(e)): every stage is per-pixel or per-sampled-row except the per-class prototype mean, which needs the
global (feature sum, count) of every class.  Each rank runs the op on its own images with
``process_group=...``; the op all-reduces one ``[C, D+1]`` fp64 buffer (C*(D+1)*8 bytes) and derives the list of
valid classes from the global counts, so all ranks run the same LOOP-2 positions.  Anchors, negatives, memory
bank and the loss value stay rank-local (gradient averaging is the outer DDP's job).
"""
from __future__ import annotations

from typing import Dict

import torch


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced slice of ``n`` items for ``rank`` (first ``n % world`` ranks get one more)."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(x: Dict[str, torch.Tensor], n_lab: int, rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Slice a full batch laid out like the trainers' (labelled images first) into rank ``rank``'s share:
    labelled and unlabelled images are both split evenly, so every rank keeps the labelled-first layout."""
    B = x["rep"].shape[0]
    n_unlab = B - n_lab
    l0, l1 = shard_range(n_lab, rank, world)
    u0, u1 = shard_range(n_unlab, rank, world)
    rows = list(range(l0, l1)) + list(range(n_lab + u0, n_lab + u1))
    idx = torch.tensor(rows, dtype=torch.long, device=x["rep"].device)
    out = {k: x[k].index_select(0, idx) for k in ("rep", "rep_teacher", "low_mask", "high_mask") if k in x}
    if "labels" in x:
        out["labels"] = x["labels"].index_select(0, idx)
    out["label_l"], out["prob_l"] = x["label_l"][l0:l1], x["prob_l"][l0:l1]
    out["label_u"], out["prob_u"] = x["label_u"][u0:u1], x["prob_u"][u0:u1]
    return {k: v.contiguous() for k, v in out.items()}


def allreduce_sum_hook(group=None):
    """``proto_sum_hook`` for the oracle's multi-GPU restatement: sum the [C, D+1] buffer over ``group``."""
    import torch.distributed as dist

    def hook(local: torch.Tensor) -> torch.Tensor:
        buf = local.clone()
        dist.all_reduce(buf, group=group)
        return buf
    return hook
