"""CPU, world_size 2 over gloo: the N>1 path's host logic (batch sharding, the single all-reduce of the
per-class prototype sums, globally agreed valid classes) on the oracle's multi-GPU restatement
(SURVEY.md section 8(e))."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from arco_b200.sharded import allreduce_sum_hook, shard_batch, shard_range
from arco_b200.synth import CaseSpec, exact_case, make_bank

SPEC = CaseSpec("dist", 2, 2, 4, (16, 16), 8, queries=8, negatives=4, bank_init="fill:30", caps=[50, 40, 40, 40],
                label_mode="absent:3", seed=31)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = exact_case(SPEC, 0)
        mine = shard_batch(x, SPEC.n_lab, rank, world)
        bank, ptr, caps = make_bank(SPEC)
        torch.manual_seed(100 + rank)
        rep = mine["rep"].clone().requires_grad_(True)
        res = oracle.contra_memobank_loss(
            rep, mine["label_l"], mine["label_u"], mine["prob_l"], mine["prob_u"], mine["low_mask"], mine["high_mask"],
            bank, ptr, caps, mine["rep_teacher"], delta_n=SPEC.delta_n, num_queries=SPEC.queries,
            num_negatives=SPEC.negatives, temp=SPEC.temp, proto_sum_hook=allreduce_sum_hook())
        res.loss.backward()
        out[rank] = dict(proto=res.proto.clone(), valid=list(res.valid_classes), loss=float(res.loss.detach()),
                         grad_ok=bool(torch.isfinite(rep.grad).all()), n_img=rep.shape[0])
    finally:
        dist.destroy_process_group()


def test_shard_range_is_a_partition():
    for n in (1, 2, 5, 8, 13):
        for w in (1, 2, 3, 8):
            cuts = [shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))


def test_two_ranks_share_global_prototypes():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        results = dict(out)
    # single-process reference: prototypes and valid classes of the WHOLE batch
    x = exact_case(SPEC, 0)
    bank, ptr, caps = make_bank(SPEC)
    full = oracle.contra_memobank_loss(
        x["rep"].clone().requires_grad_(True), x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"],
        x["high_mask"], bank, ptr, caps, x["rep_teacher"], delta_n=SPEC.delta_n, num_queries=SPEC.queries,
        num_negatives=SPEC.negatives, temp=SPEC.temp)
    ok = torch.tensor([c > 0 for c in full.low_valid_counts])
    for rank in range(world):
        r = results[rank]
        assert r["n_img"] == 2 and r["grad_ok"] and r["loss"] == r["loss"]
        assert r["valid"] == full.valid_classes                      # same LOOP-2 positions on every rank
        assert torch.allclose(r["proto"][ok], full.proto[ok], rtol=1e-5, atol=1e-6)
