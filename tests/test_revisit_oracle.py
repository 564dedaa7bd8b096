"""CPU: the revisiting-loss / pool-queue oracle against golden vectors made by executing the reference trainer's own
function source (tests/golden/make_golden_step.py; train_arco_2d.py:126-136, :109-120, :400-402)."""
import numpy as np
import pytest
import torch

import oracle
from cases import REVISIT_CASES, revisit_inputs
from util import load_golden


@pytest.mark.parametrize("case", REVISIT_CASES, ids=lambda c: c["name"])
def test_oracle_matches_reference(case):
    gold = load_golden(case["name"])
    x = revisit_inputs(case)
    pool = x["pool"].clone()
    ptr = torch.zeros(1, dtype=torch.long)
    for step in range(case["steps"]):
        loss, nn_index, _, _ = oracle.revisiting_loss(pool, x["rep_u"][step], x["rep_u_teacher"][step], topk=case["topk"])
        oracle.pool_enqueue(x["rep_u_teacher"][step], pool, ptr)
        assert abs(float(loss) - float(gold[f"s{step}_loss"])) <= 1e-6 * abs(float(gold[f"s{step}_loss"]))
        assert int(ptr) == int(gold[f"s{step}_ptr"][0])
    assert np.allclose(pool.numpy(), gold["pool_after"], rtol=0, atol=1e-7)
