"""CPU restatement of ARCO's stratified pixel/voxel contrastive loss -- TEST INFRASTRUCTURE ONLY.

Follows ``compute_contra_memobank_loss`` in ``/root/reference/code/loss_helper_3d.py:271-513``
(2-D images, rep ``[B,D,H,W]``) and its 3-D twin ``/root/reference/code/loss_helper.py:442-686``
(rep ``[B,D,H,W,Z]``); the two differ only in the channels-last permutes, so this restatement
flattens the spatial axes to ``S`` once and has a single code path.  ``dequeue_and_enqueue``
is ``loss_helper_3d.py:12-32``; ``label_onehot`` is the trainer-local override
``/root/reference/code/train_arco_2d.py:492-498`` (``train_arco_3d.py:463-469``).

Written with torch CPU ops in the same op mix as the reference (sort over classes, boolean
mask gathers, mean, cosine_similarity, cross_entropy, autograd) so that, timed on host cores,
it is a fair "port" baseline.  Every reference quirk listed in SURVEY.md section 8(a) "parity
traps" is reproduced and marked ``# trap N`` below.

Tie rule: the reference sorts class probabilities with an unstable sort, so the order of
exactly-equal probabilities is implementation-defined (CPU: stable for C<=16, arbitrary above).
The oracle pins it to *stable descending* (lower class id first), which is what the CUDA
kernel implements as ``rank_i = #{j: p_j > p_i} + #{j < i: p_j == p_i}``.

Pinned by ``tests/golden/*.npz`` (outputs of the real reference, see make_golden.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import torch
import torch.nn.functional as F

DELTA_P = 0.3                 # loss_helper_3d.py:316  current_class_threshold
LOW_RANK, HIGH_RANK = 3, 20   # loss_helper_3d.py:318


def label_onehot(labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """int labels ``[B,*S]`` -> float32 one-hot ``[B,C,*S]``; ignore label -1 lands in class 0
    (trap 4) exactly as ``relu`` + ``scatter_`` does in train_arco_2d.py:492-498."""
    idx = labels.clamp_min(0).to(torch.int64).cpu().unsqueeze(1)
    out = torch.zeros((labels.shape[0], num_classes) + tuple(labels.shape[1:]), dtype=torch.float32)
    return out.scatter_(1, idx, 1.0)


def fifo_enqueue(keys: torch.Tensor, slot: list, ptr_cell: torch.Tensor, capacity: int) -> int:
    """Append ``keys`` to ``slot[0]``, keep the newest ``capacity`` rows, update the pointer with
    the reference formula (loss_helper_3d.py:12-32).  Returns the number of appended keys."""
    rows = keys.detach().to("cpu")
    merged = torch.cat((slot[0], rows), dim=0)
    if merged.shape[0] >= capacity:
        slot[0] = merged[merged.shape[0] - capacity:, :]
        new_ptr = capacity
    else:
        slot[0] = merged
        new_ptr = (int(ptr_cell) + rows.shape[0]) % capacity
    ptr_cell[0] = new_ptr
    return rows.shape[0]


@dataclass
class PixelClasses:
    """Per-class boolean maps ``[C, B*S]`` in raster order (b major, then flattened space)."""
    low_valid: torch.Tensor     # label_c * low_mask != 0                      (:341,:365)
    anchor: torch.Tensor        # low_valid & prob_c > 0.3                     (:369-371)
    key: torch.Tensor           # hard-negative & rank test                    (:372-374,:388-401)
    low_valid_sum: torch.Tensor  # float sum the reference tests against 0     (:413)


def classify_pixels(label_l, label_u, prob_l, prob_u, low_mask, high_mask, delta_n: float) -> PixelClasses:
    """Sub-system 1 (a1-a3): masks, thresholds and the teacher-rank test, per class."""
    n_lab, C = label_l.shape[0], label_l.shape[1]
    lab = torch.cat((label_l, label_u), dim=0).flatten(2)           # [B,C,S] int64
    B = lab.shape[0]
    prob = torch.cat((prob_l, prob_u), dim=0).flatten(2)            # [B,C,S] f32
    low = lab * low_mask.flatten(2)                                 # int64 * f32 -> f32 (:341)
    high = lab * high_mask.flatten(2)                               # (:342)

    # rank of every class under a stable descending sort (pinned tie rule, see module docstring)
    order = torch.sort(prob, dim=1, descending=True, stable=True).indices   # [B,C,S]
    rank = torch.empty_like(order)
    rank.scatter_(1, order, torch.arange(C, device=prob.device).view(1, C, 1).expand_as(order))

    in_top = rank < LOW_RANK                                        # class sits in ranks [0,3)   (:395)
    in_mid = (rank >= LOW_RANK) & (rank < HIGH_RANK)                # class sits in ranks [3,20)  (:388-390)
    is_lab = torch.arange(B, device=prob.device).view(B, 1, 1) < n_lab
    # labelled images: top-3 AND the pixel is NOT labelled with this class (:397-399) -- combined with
    # the hard mask (which needs the label) this is always empty (trap 3); kept for fidelity.
    class_mask = torch.where(is_lab, in_top & (lab == 0), in_mid)

    low_valid = low != 0
    anchor = (prob > DELTA_P) & low_valid
    hard = (prob < delta_n) & (high != 0)
    key = hard & class_mask

    def cmajor(x):  # [B,C,S] -> [C, B*S] raster order
        return x.permute(1, 0, 2).reshape(C, -1)

    return PixelClasses(cmajor(low_valid), cmajor(anchor), cmajor(key), low.sum(dim=(0, 2)))


@dataclass
class OracleResult:
    new_keys: List[int]
    loss: torch.Tensor
    prototype: Optional[torch.Tensor] = None          # only with momentum_prototype
    # intermediates for stage-level parity checks
    classes: Optional[PixelClasses] = None
    anchor_lists: List[torch.Tensor] = field(default_factory=list)   # per class: flat pixel ids, raster order
    key_lists: List[torch.Tensor] = field(default_factory=list)
    low_valid_counts: List[int] = field(default_factory=list)
    proto: Optional[torch.Tensor] = None              # [C,D] (NaN rows for absent classes)
    valid_classes: List[int] = field(default_factory=list)
    slots: List[dict] = field(default_factory=list)   # per LOOP-2 position: dict(active, bank_class, idx_a, idx_n, logits)


def contra_memobank_loss(
    rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask,
    memobank, queue_ptrlis, queue_size, rep_teacher,
    momentum_prototype=None, i_iter=0, delta_n=1.0,
    sampler: Optional[Callable[[int, int], torch.Tensor]] = None,
    num_queries=256, num_negatives=512, temp=0.5,
    proto_sum_hook: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
) -> OracleResult:
    """The whole hot path.  ``sampler(high, shape) -> int64[shape]`` replaces the reference's
    ``function_to_use`` (:327-338); tests pass a replay function so both sides see identical
    indices.  ``proto_sum_hook`` (multi-GPU restatement) receives the local
    ``[C, D+1]`` float64 (feature sums, low-valid counts) and returns the global one."""
    if sampler is None:
        sampler = lambda high, shape: torch.randint(high, (shape,))
    C = label_l.shape[1]
    D = rep.shape[1]
    B = rep.shape[0]
    # channels-last rows, raster order (b, s): the reference indexes a permuted view (:344-345, :376-377)
    rows_s = rep.flatten(2).permute(0, 2, 1).reshape(-1, D)                  # keeps autograd graph
    rows_t = rep_teacher.detach().flatten(2).permute(0, 2, 1).reshape(-1, D)

    px = classify_pixels(label_l, label_u, prob_l, prob_u, low_mask, high_mask, delta_n)
    res = OracleResult(new_keys=[], loss=None, classes=px)

    anchor_feats, protos = [], []
    local = torch.zeros(C, D + 1, dtype=torch.float64, device=rep.device)
    for c in range(C):
        a_idx = torch.nonzero(px.anchor[c]).flatten()
        k_idx = torch.nonzero(px.key[c]).flatten()
        res.anchor_lists.append(a_idx)
        res.key_lists.append(k_idx)
        anchor_feats.append(rows_s[a_idx])                                   # (:377) differentiable gather
        members = rows_t[px.low_valid[c]]                                    # (:382)
        if proto_sum_hook is None:
            protos.append(members.mean(dim=0, keepdim=True))                 # (:380-384) NaN row if absent
        else:
            local[c, :D] = members.double().sum(dim=0)
            local[c, D] = members.shape[0]
        res.new_keys.append(fifo_enqueue(rows_t[k_idx], memobank[c], queue_ptrlis[c], queue_size[c]))  # (:403-411)
        res.low_valid_counts.append(int(px.low_valid[c].sum()))
    if proto_sum_hook is None:
        present = [c for c in range(C) if float(px.low_valid_sum[c]) > 0]    # (:413-415)
    else:
        glob = proto_sum_hook(local)
        protos = [(glob[c, :D] / glob[c, D]).float().unsqueeze(0) for c in range(C)]
        present = [c for c in range(C) if float(glob[c, D]) > 0]
    res.valid_classes = present
    res.proto = torch.cat(protos)

    if len(present) <= 1:                                                    # trap 5 (:417-424)
        res.loss = torch.tensor(0.0) * rep.sum()
        if momentum_prototype is not None:
            res.prototype = momentum_prototype
        return res

    n_slots = len(present)
    total = torch.tensor(0.0, device=rep.device)
    proto_out = torch.zeros(C, num_queries, 1, D, device=rep.device)
    for pos in range(n_slots):
        # trap 1: anchors and prototype are addressed by loop POSITION, the bank by CLASS ID (:437-438,:455,:466,:481)
        bank_class = present[pos]
        bank = memobank[bank_class][0]
        slot = dict(active=False, bank_class=bank_class, idx_a=None, idx_n=None, logits=None)
        res.slots.append(slot)
        n_anchor = anchor_feats[pos].shape[0]
        if n_anchor == 0 or bank.shape[0] == 0:
            total = total + 0 * rep.sum()                                    # trap 2: still counted in the divisor
            continue
        slot["active"] = True
        idx_a = sampler(n_anchor, num_queries)
        queries = anchor_feats[pos][idx_a.to(rep.device)]                    # [Q,D] trap 8: duplicates allowed
        with torch.no_grad():
            idx_n = sampler(bank.shape[0], num_queries * num_negatives)
            negatives = bank[idx_n].reshape(num_queries, num_negatives, D).to(queries.device)   # .cuda() at :466
            positive = res.proto[pos].view(1, 1, D).repeat(num_queries, 1, 1)
            if momentum_prototype is not None:                               # a11 (:488-497)
                if not (momentum_prototype == 0).all():
                    decay = min(1 - 1 / i_iter, 0.999)
                    positive = (1 - decay) * positive + decay * momentum_prototype[bank_class]
                proto_out[bank_class] = positive.clone()
            keys = torch.cat((positive, negatives), dim=1)                   # [Q,1+N,D]
        logits = torch.cosine_similarity(queries.unsqueeze(1), keys, dim=2)  # (:503-505)
        total = total + F.cross_entropy(logits / temp, torch.zeros(num_queries, dtype=torch.long, device=logits.device))
        slot.update(idx_a=idx_a, idx_n=idx_n, logits=logits.detach())
    res.loss = total / n_slots                                               # (:511)
    if momentum_prototype is not None:
        res.prototype = proto_out
    return res
