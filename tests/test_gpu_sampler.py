"""GPU: the in-kernel Philox/Feistel sampler against the STRUCTURE of the reference samplers.

The random streams differ by construction (torch CPU MT19937 vs Philox counters), so parity is on what
the reference fixes deterministically (SURVEY.md section 8(a) "Sampler semantics"): range, which path runs,
draws per grid block / per 1-D stratum, antithetic mirror pairs, number of pads, and -- statistically --
uniformity inside a block and of the final shuffle."""
import numpy as np
import pytest
import torch

import oracle
from oracle.samplers_oracle import sampler_plan

pytestmark = pytest.mark.gpu

FUNCS = {"smc": 1, "asmc": 2, "rand": 0}


def _draw(func, high, shape, seed=7, stream_id=3):
    from arco_b200.samplers import _sample
    return _sample(FUNCS[func], high, shape, device=torch.device("cuda", 0), seed=seed, stream_id=stream_id).cpu().numpy()


CASES = [(1, 256), (5, 16), (15, 256), (16, 256), (17, 31), (40, 7), (56, 256), (57, 256), (64, 256), (100, 256),
         (1000, 256), (6133, 256), (300, 2048), (29929, 4096), (30000, 8192), (30000, 131072), (50000, 131072),
         (3000, 1), (5000, 20), (4014080, 256), (250000, 256)]


@pytest.mark.parametrize("func", ["smc", "asmc", "rand"])
@pytest.mark.parametrize("high,shape", CASES)
def test_structure(func, high, shape):
    out = _draw(func, high, shape)
    assert out.shape == (shape,)
    assert out.min() >= 0 and out.max() < high
    plan = sampler_plan(high, shape, func)
    if plan.path == "grid":
        blk = plan.block_of(out)
        counts = np.bincount(blk[blk >= 0], minlength=16)
        # every block contributes at most per_block draws; together with the pads they fill `shape`
        n_grid_cells = plan.edge * plan.edge
        if n_grid_cells <= high:
            # no draw can be dropped: the first min(16*per_block, shape) entries are the shuffled draws
            n_keep = min(plan.n_structured, shape)
            head = plan.block_of(out[:n_keep])
            assert (head >= 0).all()
            if plan.n_structured <= shape:
                assert (np.bincount(head, minlength=16) == plan.per_block).all()
        else:
            assert (counts[:12] <= plan.per_block + (shape - plan.n_structured if shape > plan.n_structured else 0) + shape).all()
            # blocks 0..11 never lose a draw; if nothing is truncated they hold exactly per_block in the head
            n_drop_max = 4 * plan.per_block
            assert plan.n_structured - n_drop_max <= shape + n_drop_max
    elif plan.path == "strata":
        n = plan.n_structured
        # the shuffle covers pads too, so only totals are fixed: each stratum holds >= per_stratum entries
        strata = out[out < plan.strata * 16] // 16
        counts = np.bincount(strata, minlength=plan.strata)
        assert (counts >= plan.per_stratum).all()
        assert counts.sum() - n <= shape - n


@pytest.mark.parametrize("high,shape", [(100, 256), (6133, 256), (30000, 8192), (10000, 4096)])
def test_antithetic_pairs(high, shape):
    """Every structured draw of 'asmc' has its point mirror (center - x) in the output (edge^2 <= high cases)."""
    plan = sampler_plan(high, shape, "asmc")
    assert plan.path == "grid" and plan.edge ** 2 <= high and plan.n_structured <= shape
    out = _draw("asmc", high, shape)[: plan.n_structured]
    step = plan.edge // 4
    blk = plan.block_of(out)
    for k in range(16):
        bi, bj = divmod(k, 4)
        r0, c0 = bi * step, bj * step
        nr = plan.edge - r0 if bi == 3 else step
        nc = plan.edge - c0 if bj == 3 else step
        center = (2 * r0 + nr - 1) * plan.edge + (2 * c0 + nc - 1)
        vals = np.sort(out[blk == k])
        assert np.array_equal(vals, np.sort(center - vals)), f"block {k} is not mirror symmetric"


def test_seed_reproducible_and_streams_differ():
    a = _draw("smc", 30000, 131072, seed=11, stream_id=5)
    b = _draw("smc", 30000, 131072, seed=11, stream_id=5)
    c = _draw("smc", 30000, 131072, seed=11, stream_id=6)
    d = _draw("smc", 30000, 131072, seed=12, stream_id=5)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, c) and not np.array_equal(a, d)


def test_drop_and_pad_counts():
    """edge^2 < high: exactly shape - 16*per_block pads.  edge^2 > high: draws landing on the cells >= high of
    the last image row are dropped (Binomial count), everything else survives, pads fill up to `shape`."""
    high, shape = 30000, 131072            # edge 173, edge^2 = 29929 < high
    plan = sampler_plan(high, shape, "smc")
    out = _draw("smc", high, shape)
    assert shape - plan.n_structured == 320
    head = plan.block_of(out[: plan.n_structured])
    assert (np.bincount(head, minlength=16) == plan.per_block).all()

    high = 50000                            # edge 224, edge^2 = 50176 > high: 176 cells of the last row are out
    plan = sampler_plan(high, shape, "smc")
    assert plan.edge == 224 and plan.per_block == 8220
    step = 56
    p_drop = np.array([8, 56, 56, 56]) / float(step * step)
    drops = []
    for sid in range(8):
        out = _draw("smc", high, shape, stream_id=100 + sid)
        assert out.max() < high
        counts = np.bincount(plan.block_of(out), minlength=16)
        slack = 80                                                   # pads (uniform over [0,high)) land in blocks too
        # E[survivors] ~= shape here, so a few entries are either truncated away or padded in
        assert (counts[:12] >= plan.per_block - slack).all() and (counts[:12] <= plan.per_block + slack).all(), counts
        exp_bottom = plan.per_block * (1 - p_drop)
        sig = np.sqrt(plan.per_block * p_drop * (1 - p_drop))
        assert (np.abs(counts[12:] - exp_bottom) < 6 * sig + slack).all(), counts[12:]
        drops.append(4 * plan.per_block - counts[12:].sum())
    exp = plan.per_block * p_drop.sum()
    assert abs(np.mean(drops) - exp) < 40, (np.mean(drops), exp)


def test_uniform_within_block_and_shuffle():
    high, shape = 40000, 131072            # edge 200, no drops, no pads beyond 16*pps
    plan = sampler_plan(high, shape, "smc")
    out = _draw("smc", high, shape)[: plan.n_structured]
    # chi-square of cell occupancy inside block 5 (50x50 cells)
    blk = plan.block_of(out)
    cells = out[blk == 5]
    _, cnt = np.unique(cells, return_counts=True)
    n, k = len(cells), 2500
    full = np.concatenate([cnt, np.zeros(k - len(cnt))])
    chi2 = ((full - n / k) ** 2 / (n / k)).sum()
    assert abs(chi2 - k) < 6 * np.sqrt(2 * k), chi2
    # the shuffle: block ids along the sequence look iid -- each quarter of the output holds ~1/4 of every block
    q = len(out) // 4
    for part in range(4):
        c = np.bincount(plan.block_of(out[part * q:(part + 1) * q]), minlength=16)
        assert np.all(np.abs(c - plan.per_block / 4) < 6 * np.sqrt(plan.per_block / 4))


def test_dropin_sampler_names():
    import arco_b200
    for fn in (arco_b200.grid_monte_carlo_sample, arco_b200.grid_as_monte_carlo_sample):
        out = fn(6133, 256)
        assert out.dtype == torch.int64 and out.shape == (256,) and out.is_cuda
        assert int(out.min()) >= 0 and int(out.max()) < 6133
    out = arco_b200.monte_carlo_sample(40, 64)
    assert out.shape == (64,) and int(out.max()) < 40
