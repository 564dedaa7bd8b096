"""CPU oracle for the ARCO stratified contrastive loss -- TEST INFRASTRUCTURE ONLY.

Nothing under ``arco_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.

Parity status: PINNED.  The reference ships no tests or golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself, run in the
build container by ``tests/golden/make_golden.py`` (imports
``/root/reference/code/loss_helper_3d.py`` and ``loss_helper.py`` unmodified) and
committed under ``tests/golden/``.
"""
from .contra_oracle import (  # noqa: F401
    OracleResult,
    classify_pixels,
    contra_memobank_loss,
    fifo_enqueue,
    label_onehot,
)
from .samplers_oracle import (  # noqa: F401
    SamplerPlan,
    antithetic_strata_sample,
    grid_antithetic_sample,
    grid_strata_sample,
    sampler_plan,
    strata_sample,
)
from .revisit_oracle import pool_enqueue, revisiting_loss  # noqa: F401,E402
from .stepterms_oracle import equivariance_loss, tps_grid, unsupervised_loss, warp  # noqa: F401,E402
from . import producers_oracle  # noqa: F401,E402
