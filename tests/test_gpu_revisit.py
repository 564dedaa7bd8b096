"""GPU: arco_b200.get_revisiting_loss / revisit_enqueue (one streaming CUDA pass + one scale-and-copy pass) against the
reference's golden vectors and against the oracle at the trainers' full size (bs 12, K 36, D*H*W = 32.5 M)."""
import gc

import numpy as np
import pytest
import torch

import oracle
from cases import REVISIT_CASES, revisit_inputs
from util import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", REVISIT_CASES, ids=lambda c: c["name"])
def test_against_reference_golden(case):
    import arco_b200
    dev = torch.device("cuda", 0)
    gold = load_golden(case["name"])
    x = revisit_inputs(case)
    tdt = torch.bfloat16 if case["dtype"] == "bf16" else torch.float32
    pool = x["pool"].to(dev).contiguous()
    ptr = torch.zeros(1, dtype=torch.long)
    pool_o = x["pool"].clone()
    ptr_o = torch.zeros(1, dtype=torch.long)
    for step in range(case["steps"]):
        rs, rt = x["rep_u"][step].to(dev, tdt), x["rep_u_teacher"][step].to(dev, tdt)
        loss, nn_index = arco_b200.get_revisiting_loss(pool, rs, rt, topk=case["topk"], return_index=True)
        arco_b200.revisit_enqueue(rt, pool, ptr)
        torch.cuda.synchronize()
        want = float(gold[f"s{step}_loss"])
        assert abs(float(loss) - want) <= 1e-5 * abs(want), (step, float(loss), want)
        assert int(ptr) == int(gold[f"s{step}_ptr"][0])
        _, idx_o, dist_t, _ = oracle.revisiting_loss(pool_o, x["rep_u"][step], x["rep_u_teacher"][step], topk=case["topk"])
        oracle.pool_enqueue(x["rep_u_teacher"][step], pool_o, ptr_o)
        # the neighbours: same SET per row unless two student distances are closer than fp32 noise
        for b in range(case["bs"]):
            got, ref = set(nn_index[b].tolist()), set(idx_o[b].tolist())
            if got != ref:
                gap = dist_t[b].sort().values
                assert float((gap[case["topk"]] - gap[case["topk"] - 1]).abs()) < 1e-5, (step, b, got, ref)
    assert np.allclose(pool.cpu().numpy(), gold["pool_after"], rtol=0, atol=2e-7)
    # rows stay unit vectors
    assert float((pool.norm(dim=1) - 1).abs().max()) < 1e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
def test_trainer_size_against_oracle_ops(dtype):
    """bs = 12 unlabelled images, D = 496, 256 x 256 (L = 32 505 856), K = 36, topk = 5 (train_arco_2d.py:66,72,156)."""
    import arco_b200
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    bs, K, D, H, W = 12, 36, 496, 256, 256
    pool = torch.randn(K, D * H * W, device=dev, generator=g)
    pool = torch.nn.functional.normalize(pool, dim=1)
    rep_s = torch.randn(bs, D, H, W, device=dev, generator=g).to(dtype)
    # teachers near the students, both pulled towards pool rows so that the top-k is decided by real margins
    rep_t = (rep_s.float() + 0.3 * torch.randn(bs, D, H, W, device=dev, generator=g)).to(dtype)
    for b in range(bs):
        pull = pool[(3 * b) % K].view(D, H, W) * (D * H * W) ** 0.5 * 0.3
        rep_s[b] = (rep_s[b].float() + pull).to(dtype)
        rep_t[b] = (rep_t[b].float() + pull).to(dtype)
    ptr = torch.zeros(1, dtype=torch.long)
    loss, nn_index = arco_b200.get_revisiting_loss(pool, rep_s, rep_t, topk=5, return_index=True)
    torch.cuda.synchronize()
    lo, idx_o, dist_t, dist_q = oracle.revisiting_loss(pool, rep_s, rep_t, topk=5)
    assert abs(float(loss) - float(lo)) <= 1e-5 * abs(float(lo)), (float(loss), float(lo))
    assert torch.equal(nn_index[:, 0].cpu(), idx_o[:, 0].cpu())              # the pulled row is the clear nearest neighbour
    before = pool[:bs].clone()
    arco_b200.revisit_enqueue(rep_t, pool, ptr)
    torch.cuda.synchronize()
    assert int(ptr) == bs
    want = torch.nn.functional.normalize(rep_t.reshape(bs, -1).float(), dim=-1)
    assert float((pool[:bs] - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert not torch.equal(before, pool[:bs])
    del pool, rep_s, rep_t, want, before
    gc.collect()
    torch.cuda.empty_cache()


def test_shape_limits_raise():
    import arco_b200
    dev = torch.device("cuda", 0)
    pool = torch.nn.functional.normalize(torch.randn(80, 64, device=dev), dim=1)
    with pytest.raises(arco_b200.ArcoError):
        arco_b200.get_revisiting_loss(pool, torch.randn(40, 64, device=dev), torch.randn(40, 64, device=dev))
    with pytest.raises(RuntimeError):
        arco_b200.get_revisiting_loss(pool.cpu(), torch.randn(4, 64), torch.randn(4, 64))
