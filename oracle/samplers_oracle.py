"""CPU restatement of ARCO's four index samplers -- TEST INFRASTRUCTURE ONLY.

Reference (2-D file; ``loss_helper.py:206-439`` is a byte-identical twin):

* ``grid_monte_carlo_sample``     ``/root/reference/code/loss_helper_3d.py:120-184``  -> :func:`grid_strata_sample`
* ``grid_as_monte_carlo_sample``  ``loss_helper_3d.py:187-268``                       -> :func:`grid_antithetic_sample`
* ``monte_carlo_sample``          ``loss_helper_3d.py:83-117``                        -> :func:`strata_sample`
* ``as_monte_carlo_sample``       ``loss_helper_3d.py:35-80``                         -> :func:`antithetic_strata_sample`

The restatement consumes the host RNG streams (torch CPU generator, Python ``random``) in
exactly the order the reference does, so under the same seeds it returns bit-identical
index tensors; ``tests/golden/samplers.npz`` pins that.  The reference reaches its
fallbacks through a bare ``except``; probing it over ``high in [1, 50000]`` shows the
exception fires exactly when ``round(sqrt(high)) < 8`` (a 0- or 1-cell block makes a NumPy
index collapse to a 0-d array), so the fallback condition is written out explicitly here.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass

import numpy as np
import torch

CUT = 4          # 4x4 grid of blocks            (loss_helper_3d.py:121 default cut_count)
PATCH = 16       # 1-D stratum width, fallbacks  (loss_helper_3d.py:83 default patch)
MIN_GRID_EDGE = 8


def _edge_of(high: int) -> int:
    # loss_helper_3d.py:136 -- Python round() of a double; sqrt of an integer is never x.5
    return round(math.sqrt(high))


def _block_cells(edge: int):
    """Row-major ``edge x edge`` image cut into CUT x CUT blocks; the last block row / column
    absorbs the remainder (loss_helper_3d.py:141-153).  Returns 16 flat int64 arrays."""
    step = edge // CUT
    cuts = [k * step for k in range(CUT)] + [edge]
    image = np.arange(edge * edge, dtype=np.int64).reshape(edge, edge)
    return [image[cuts[r]:cuts[r + 1], cuts[c]:cuts[c + 1]].reshape(-1)
            for r in range(CUT) for c in range(CUT)]


def _finish_grid(values: torch.Tensor, high: int, shape: int) -> torch.Tensor:
    """Drop out-of-range draws, shuffle survivors, pad one uniform draw at a time, truncate
    (loss_helper_3d.py:165-182)."""
    survivors = values[values < high]
    survivors = survivors[torch.randperm(survivors.numel())]
    pads = []
    while survivors.numel() + len(pads) < shape:
        pads.append(torch.randint(high, (1, 1)).reshape(1))
    if pads:
        survivors = torch.cat([survivors] + pads)
    return survivors[:shape]


def grid_strata_sample(high: int, shape: int) -> torch.Tensor:
    """'smc' sampler.  ``per_block = shape*edge^2 // high // 16`` draws with replacement from
    every block; indices take a float32 round trip (loss_helper_3d.py:163)."""
    edge = _edge_of(high)
    if edge < MIN_GRID_EDGE:
        return strata_sample(high, shape)
    per_block = shape * edge * edge // high // (CUT * CUT)
    rows = []
    for cells in _block_cells(edge):
        shuffled = cells[torch.randperm(cells.shape[0]).numpy()]
        pick = torch.randint(cells.shape[0], (per_block,)).numpy()
        rows.append(shuffled[pick])
    values = torch.from_numpy(np.stack(rows).astype(np.float32)).flatten().long()
    return _finish_grid(values, high, shape)


def grid_antithetic_sample(high: int, shape: int) -> torch.Tensor:
    """'asmc' sampler.  ``per_block // 2`` draws per block plus their point mirror
    ``center - x`` with ``center = int(2*mean(block))`` (loss_helper_3d.py:224-238)."""
    edge = _edge_of(high)
    if edge < MIN_GRID_EDGE:
        return antithetic_strata_sample(high, shape)
    half = (shape * edge * edge // high // (CUT * CUT)) // 2
    rows = []
    for cells in _block_cells(edge):
        center = np.int64(2 * np.mean(cells))
        shuffled = cells[torch.randperm(cells.shape[0]).numpy()]
        pick = torch.randint(cells.shape[0], (half,)).numpy()
        drawn = shuffled[pick]
        rows.append(drawn)
        rows.append(center - drawn)
    values = torch.from_numpy(np.stack(rows).astype(np.float32)).flatten().long()
    return _finish_grid(values, high, shape)


def _strata(high: int, shape: int, antithetic: bool) -> torch.Tensor:
    # loss_helper_3d.py:84-85 / :36-37 -- too few samples for the strata, or fewer than one stratum
    if high // PATCH > shape or high < PATCH:
        return torch.randint(high, size=(shape,))
    strata = high // PATCH
    per_stratum = shape // strata
    # (the reference's `blocks > shape` arm, :92-96 / :44-48, is unreachable after the test above)
    values = []
    for k in range(strata):
        lo, hi = k * PATCH, (k + 1) * PATCH - 1
        if antithetic:
            drawn = [random.randint(lo, hi) for _ in range(per_stratum // 2)]
            values.extend(drawn)
            values.extend((lo + hi) - x for x in drawn)          # (2k+1)*PATCH - 1 - x
        else:
            values.extend(random.randint(lo, hi) for _ in range(per_stratum))
    while len(values) < shape:
        values.append(random.randint(0, high - 1))
    out = torch.tensor(values, dtype=torch.float32).reshape(shape).long()
    return out[torch.randperm(shape)]


def strata_sample(high: int, shape: int) -> torch.Tensor:
    """1-D stratified fallback of 'smc' (loss_helper_3d.py:83-117)."""
    return _strata(high, shape, antithetic=False)


def antithetic_strata_sample(high: int, shape: int) -> torch.Tensor:
    """1-D antithetic fallback of 'asmc' (loss_helper_3d.py:35-80)."""
    return _strata(high, shape, antithetic=True)


# --------------------------------------------------------------------------------------
# Deterministic structure of a sampler call -- what a device sampler must reproduce in
# distribution.  Used by the GPU tests to check arco_sample without sharing an RNG stream.
# --------------------------------------------------------------------------------------
@dataclass
class SamplerPlan:
    path: str            # "grid" | "strata" | "uniform"
    antithetic: bool
    high: int
    shape: int
    edge: int = 0        # grid: image edge
    per_block: int = 0   # grid: draws per block (already halved+mirrored for antithetic -> 2*(pps//2))
    strata: int = 0      # strata: number of 16-wide strata
    per_stratum: int = 0 # strata: structured draws per stratum
    n_structured: int = 0  # draws made before drop / pad

    def block_of(self, values: np.ndarray) -> np.ndarray:
        """Grid block id (0..15) of each index < edge^2, -1 for indices only a pad can produce."""
        step = self.edge // CUT
        r = np.minimum(values // self.edge // step, CUT - 1)
        c = np.minimum(values % self.edge // step, CUT - 1)
        blk = r * CUT + c
        return np.where(values < self.edge * self.edge, blk, -1)


def sampler_plan(high: int, shape: int, func: str) -> SamplerPlan:
    if func not in ("smc", "asmc"):
        return SamplerPlan("uniform", False, high, shape)
    anti = func == "asmc"
    edge = _edge_of(high)
    if edge >= MIN_GRID_EDGE:
        pps = shape * edge * edge // high // (CUT * CUT)
        per_block = 2 * (pps // 2) if anti else pps
        return SamplerPlan("grid", anti, high, shape, edge=edge, per_block=per_block,
                           n_structured=16 * per_block)
    if high // PATCH > shape or high < PATCH:
        return SamplerPlan("uniform", anti, high, shape)
    strata = high // PATCH
    per = shape // strata
    per = 2 * (per // 2) if anti else per
    return SamplerPlan("strata", anti, high, shape, strata=strata, per_stratum=per,
                       n_structured=strata * per)
