"""Synthetic inputs for the contrastive-loss hot path (tests, golden vectors, bench).

Two generators:

* :func:`exact_case` -- NumPy ``RandomState`` integers turned into floats by exact IEEE
  operations only, so the very same bits come out on every machine.  Used for the golden
  vectors (``tests/golden``) and the parity tests.
* :func:`bench_inputs` -- device-side ``torch.randn`` / ``softmax`` data of the shapes in
  BASELINE.json / SURVEY.md section 8(d) (first half of the batch labelled, second half unlabelled,
  20 % low/high-entropy masks, 5 % ignore labels, N(0,1) features).

Shapes follow the reference call site ``/root/reference/code/train_arco_2d.py:394-398``:
``rep``/``rep_teacher`` ``[B,D,*S]``, one-hot int64 labels ``[B/2,C,*S]``, teacher probabilities
``[B/2,C,*S]`` f32, masks ``[B,1,*S]`` f32.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass
class CaseSpec:
    name: str
    n_lab: int
    n_unlab: int
    classes: int
    spatial: Tuple[int, ...]
    feat: int
    queries: int = 16
    negatives: int = 8
    func: str = "smc"
    delta_n: float = 0.97
    temp: float = 0.5
    caps: Optional[Sequence[int]] = None      # queue_size per class (default 50000 / 30000 like the trainers)
    bank_init: str = "zeros1"                 # zeros1 (2-D trainer) | randn1 (3-D trainer) | fill:<n>
    label_mode: str = "iid"                   # iid | blocky | absent:<c> | single:<c> | noanchor:<c>
    ignore_frac: float = 0.05
    mask_frac: float = 0.2
    mask_mode: str = "iid"                    # iid | coherent (one smooth entropy-like field, low = bottom / high = top mask_frac of it)
    steps: int = 1
    dtype: str = "f32"                        # f32 | bf16 (representation tensors)
    seed: int = 1337
    momentum: str = ""                        # "" (none) | "zeros" | "rand": momentum_prototype [C,Q,1,D] (a11)
    i_iter: int = 0

    @property
    def batch(self) -> int:
        return self.n_lab + self.n_unlab

    @property
    def pixels(self) -> int:
        return self.batch * int(np.prod(self.spatial))

    def queue_sizes(self) -> List[int]:
        if self.caps is not None:
            return list(self.caps)
        return [50000] + [30000] * (self.classes - 1)


def _exact_normal_like(rs: np.random.RandomState, shape) -> np.ndarray:
    """Sum of three uniform integers / 256: bell-shaped, zero-mean, exactly representable in bf16-free
    fp32 (and in bf16 after rounding is applied by the caller)."""
    a = rs.randint(-256, 257, size=shape).astype(np.int32)
    b = rs.randint(-256, 257, size=shape).astype(np.int32)
    c = rs.randint(-256, 257, size=shape).astype(np.int32)
    return ((a + b + c).astype(np.float32) / np.float32(256.0)).astype(np.float32)


def _exact_probs(rs: np.random.RandomState, batch: int, classes: int, space: int) -> np.ndarray:
    """Tie-free probability vectors from distinct powers of two (exact IEEE division)."""
    span = max(classes + 4, 12)
    expo = np.argsort(rs.random_sample((batch, span, space)), axis=1)[:, :classes, :]   # distinct ints in [0,span)
    expo = np.minimum(expo, 40)
    w = np.ldexp(1.0, expo.astype(np.int64))
    return (w / w.sum(axis=1, keepdims=True)).astype(np.float32)


def _labels(rs: np.random.RandomState, spec: CaseSpec) -> np.ndarray:
    B, C = spec.batch, spec.classes
    sp = tuple(spec.spatial)
    mode, _, arg = spec.label_mode.partition(":")
    if mode == "blocky":
        coarse = tuple(max(1, (s + 3) // 4) for s in sp)
        lab = rs.randint(0, C, size=(B,) + coarse)
        for ax, s in enumerate(sp):
            lab = np.repeat(lab, 4, axis=ax + 1).take(np.arange(s), axis=ax + 1)
    else:
        lab = rs.randint(0, C, size=(B,) + sp)
    if mode == "absent":
        c = int(arg)
        lab = np.where(lab == c, (c + 1) % C, lab)
    elif mode == "single":
        lab = np.full_like(lab, int(arg))
    lab = lab.astype(np.int64)
    if spec.ignore_frac > 0:
        ign = rs.random_sample(lab.shape) < spec.ignore_frac
        ign[: spec.n_lab] = False                     # only unlabelled pixels carry the ignore label
        lab = np.where(ign, -1, lab)
    return lab


def _coherent_masks(rs: np.random.RandomState, shape, frac: float, window: int = 15):
    """Spatially coherent low / high masks the way the trainers derive them (train_arco_2d.py:356-392): ONE entropy-like
    field per image, low = its bottom ``frac`` quantile, high = its top ``frac`` quantile (disjoint blobs instead of iid
    pixels).  The field is a box sum of integer noise and the thresholds are integer order statistics, so the masks are the
    same bits on every machine."""
    field = rs.randint(-256, 257, size=shape).astype(np.int64)
    for ax in 3 * list(range(1, len(shape))):                 # three box passes per axis ~ a Gaussian: clean blobs
        n = shape[ax]
        w = min(window, n)
        pad = [(0, 0)] * len(shape)
        pad[ax] = (w // 2, w - 1 - w // 2)
        c = np.cumsum(np.pad(field, pad, mode="wrap"), axis=ax)
        c = np.concatenate([np.zeros_like(np.take(c, [0], axis=ax)), c], axis=ax)
        field = np.take(c, np.arange(w, w + n), axis=ax) - np.take(c, np.arange(0, n), axis=ax)
    flat = np.sort(field.reshape(shape[0], -1), axis=1)
    k = max(1, int(frac * flat.shape[1]))
    expand = (slice(None),) + (None,) * (len(shape) - 1)
    lo_thr = flat[:, k - 1][expand]
    hi_thr = flat[:, flat.shape[1] - k][expand]
    return field <= lo_thr, field >= hi_thr


def onehot_relu(lab: torch.Tensor, classes: int) -> torch.Tensor:
    """What the trainers feed the loss: relu'd labels scattered to one-hot, as int64
    (train_arco_2d.py:349-350,394 ``label_onehot(...).cuda().long()``)."""
    idx = lab.clamp_min(0).unsqueeze(1)
    out = torch.zeros((lab.shape[0], classes) + tuple(lab.shape[1:]), dtype=torch.int64, device=lab.device)
    return out.scatter_(1, idx, 1)


def exact_case(spec: CaseSpec, step: int = 0) -> dict:
    """Machine-independent inputs for ``spec`` at call number ``step`` (CPU tensors)."""
    rs = np.random.RandomState(spec.seed + 7919 * step)
    B, C, D = spec.batch, spec.classes, spec.feat
    sp = tuple(spec.spatial)
    S = int(np.prod(sp))
    lab = _labels(rs, spec)
    prob = _exact_probs(rs, B, C, S).reshape((B, C) + sp)
    mode, _, arg = spec.label_mode.partition(":")
    if mode == "noanchor":
        # make class <arg> never confident on its own pixels: its probability becomes the smallest
        c = int(arg)
        tiny = prob.min(axis=1) * np.float32(0.5)
        prob[:, c] = tiny
    rep = _exact_normal_like(rs, (B, D) + sp)
    rep_t = _exact_normal_like(rs, (B, D) + sp)
    valid = (lab >= 0)
    if spec.mask_mode == "coherent":
        low, high = _coherent_masks(rs, lab.shape, spec.mask_frac)
        low &= valid
        high &= valid
    else:
        low = (rs.random_sample(lab.shape) < spec.mask_frac) & valid
        high = (rs.random_sample(lab.shape) < spec.mask_frac) & valid
    low[: spec.n_lab] = valid[: spec.n_lab]
    high[: spec.n_lab] = valid[: spec.n_lab]

    tdtype = torch.bfloat16 if spec.dtype == "bf16" else torch.float32
    lab_t = torch.from_numpy(lab)
    onehot = onehot_relu(lab_t, C)
    out = dict(
        labels=lab_t,
        rep=torch.from_numpy(rep).to(tdtype),
        rep_teacher=torch.from_numpy(rep_t).to(tdtype),
        label_l=onehot[: spec.n_lab].contiguous(),
        label_u=onehot[spec.n_lab:].contiguous(),
        prob_l=torch.from_numpy(prob[: spec.n_lab]).contiguous(),
        prob_u=torch.from_numpy(prob[spec.n_lab:]).contiguous(),
        low_mask=torch.from_numpy(low.astype(np.float32)).unsqueeze(1),
        high_mask=torch.from_numpy(high.astype(np.float32)).unsqueeze(1),
    )
    if spec.momentum:
        shape = (C, spec.queries, 1, D)
        mom = np.zeros(shape, np.float32) if spec.momentum == "zeros" else _exact_normal_like(rs, shape)
        out["momentum_prototype"] = torch.from_numpy(mom)
    return out


def make_bank(spec: CaseSpec):
    """Caller-owned memory bank exactly as the trainers build it (train_arco_2d.py:147-154,
    train_arco_3d.py:144-151): ``memobank[c] = [tensor[n,D]]`` on CPU, pointers, capacities."""
    rs = np.random.RandomState(spec.seed ^ 0x5EED)
    kind, _, arg = spec.bank_init.partition(":")
    memobank, ptrs = [], []
    caps = spec.queue_sizes()
    for c in range(spec.classes):
        if kind == "zeros1":
            rows = np.zeros((1, spec.feat), np.float32)
        elif kind == "randn1":
            rows = _exact_normal_like(rs, (1, spec.feat))
        elif kind == "fill":
            n = min(int(arg), caps[c]) if arg else caps[c]
            rows = _exact_normal_like(rs, (n, spec.feat))
        else:
            raise ValueError(spec.bank_init)
        memobank.append([torch.from_numpy(rows)])
        ptrs.append(torch.zeros(1, dtype=torch.long))
    return memobank, ptrs, caps


# --------------------------------------------------------------------------------------
# Benchmark workloads; generated on the target device.
# --------------------------------------------------------------------------------------
WORKLOADS = {
    # name: (n_lab, n_unlab, C, spatial, D, dtype)
    "acdc2d_loss":      dict(n_lab=4,  n_unlab=4,  classes=4,  spatial=(256, 256),     feat=64,  dtype="f32"),
    "acdc2d_trainstep": dict(n_lab=12, n_unlab=12, classes=4,  spatial=(256, 256),     feat=496, dtype="bf16"),
    "la3d":             dict(n_lab=2,  n_unlab=2,  classes=2,  spatial=(112, 112, 80), feat=16,  dtype="f32"),
    "cityscapes":       dict(n_lab=8,  n_unlab=8,  classes=19, spatial=(512, 512),     feat=256, dtype="f32"),
}


def _coherent_field(shape, device, g, window: int = 31):
    """Smooth entropy-like field per image (box-filtered N(0,1) noise): thresholding it at its quantiles gives blob-shaped
    masks like the entropy masks of a real prediction, instead of iid pixels."""
    import torch.nn.functional as F
    f = torch.randn((shape[0], 1) + tuple(shape[1:]), device=device, generator=g)
    nd = len(shape) - 1
    w = [min(window if nd == 2 else 9, s) | 1 for s in shape[1:]]
    w = [k if k <= s else k - 2 for k, s in zip(w, shape[1:])]
    pool = F.avg_pool2d if nd == 2 else F.avg_pool3d
    for _ in range(3):                                        # three box passes ~ a Gaussian: clean blobs, scale ~ 50 pixels
        f = pool(f, kernel_size=w, stride=1, padding=[k // 2 for k in w], count_include_pad=False)
    return f[:, 0]


def bench_inputs(name: str, device, seed: int = 1337, blocky: bool = False, n_lab=None, n_unlab=None, coherent: bool = False):
    """Device-resident synthetic batch for workload ``name``.  Returns (spec, tensors).  ``coherent``: the unlabelled images'
    low / high masks are the bottom / top 20 % of ONE smooth field per image (how the trainers derive them from the entropy,
    train_arco_2d.py:356-392) instead of two independent Bernoulli(0.2) draws per pixel."""
    cfg = dict(WORKLOADS[name])
    if n_lab is not None:
        cfg["n_lab"] = n_lab
    if n_unlab is not None:
        cfg["n_unlab"] = n_unlab
    spec = CaseSpec(name=name, queries=256, negatives=512, bank_init="fill:", seed=seed, **cfg)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    B, C, D, sp = spec.batch, spec.classes, spec.feat, tuple(spec.spatial)
    tdtype = torch.bfloat16 if spec.dtype == "bf16" else torch.float32
    if blocky:
        coarse = tuple(max(1, s // 16) for s in sp)
        lab = torch.randint(0, C, (B,) + coarse, device=device, generator=g)
        for ax in range(len(sp)):
            lab = lab.repeat_interleave(16, dim=ax + 1)
        lab = lab[(slice(None),) + tuple(slice(0, s) for s in sp)].contiguous()
    else:
        lab = torch.randint(0, C, (B,) + sp, device=device, generator=g)
    ign = torch.rand(lab.shape, device=device, generator=g) < 0.05
    ign[: spec.n_lab] = False
    lab = torch.where(ign, torch.full_like(lab, -1), lab)
    valid = lab >= 0
    if coherent:
        f = _coherent_field(lab.shape, device, g)
        flat = f.reshape(B, -1).float()
        k = max(1, int(0.2 * flat.shape[1]))
        srt = flat.sort(dim=1).values
        ex = (slice(None),) + (None,) * len(sp)
        low = (f <= srt[:, k - 1][ex]) & valid
        high = (f >= srt[:, flat.shape[1] - k][ex]) & valid
    else:
        low = (torch.rand(lab.shape, device=device, generator=g) < 0.2) & valid
        high = (torch.rand(lab.shape, device=device, generator=g) < 0.2) & valid
    low[: spec.n_lab] = valid[: spec.n_lab]
    high[: spec.n_lab] = valid[: spec.n_lab]
    prob = torch.softmax(torch.randn((B, C) + sp, device=device, generator=g), dim=1)
    onehot = onehot_relu(lab, C)
    rep = torch.randn((B, D) + sp, device=device, generator=g, dtype=torch.float32).to(tdtype)
    rep_t = torch.randn((B, D) + sp, device=device, generator=g, dtype=torch.float32).to(tdtype)
    tensors = dict(
        labels=lab, rep=rep, rep_teacher=rep_t,
        label_l=onehot[: spec.n_lab].contiguous(), label_u=onehot[spec.n_lab:].contiguous(),
        prob_l=prob[: spec.n_lab].contiguous(), prob_u=prob[spec.n_lab:].contiguous(),
        low_mask=low.float().unsqueeze(1), high_mask=high.float().unsqueeze(1),
    )
    return spec, tensors


def bench_bank(spec: CaseSpec, seed: int = 1337, cold: bool = False):
    """``cold``: the trainers' own initial bank (one zero row per class in 2-D, train_arco_2d.py:147-154; one N(0,1) row
    in 3-D, train_arco_3d.py:144-151) -- the reference-faithful state of SURVEY.md trap 3.  Otherwise:
    banks pre-filled to capacity with N(0,1) rows (CPU lists, the trainers' layout).  With a bf16 representation
    head every row a real run ever enqueues is a bf16 teacher row, so the steady-state bank is bf16-exact: the rows
    are rounded to bf16 values (still handed over as fp32 tensors, as the trainers hold them)."""
    g = torch.Generator()
    g.manual_seed(seed)
    caps = spec.queue_sizes()
    if cold:
        three_d = len(spec.spatial) == 3
        memobank = [[torch.randn(1, spec.feat, generator=g) if three_d else torch.zeros(1, spec.feat)]
                    for _ in range(spec.classes)]
        return memobank, [torch.zeros(1, dtype=torch.long) for _ in range(spec.classes)], caps
    memobank = [[torch.randn(caps[c], spec.feat, generator=g)] for c in range(spec.classes)]
    if spec.dtype == "bf16":
        memobank = [[m[0].to(torch.bfloat16).to(torch.float32)] for m in memobank]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(spec.classes)]
    return memobank, ptrs, caps
