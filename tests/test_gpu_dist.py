"""GPU, needs >= 2 devices (skipped otherwise): the batch-sharded CUDA op over NCCL against the oracle's
multi-GPU restatement (global prototypes / valid classes, rank-local anchors, negatives and banks)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scenario(scenario, rank, world, dev):
    import arco_b200
    import oracle
    from arco_b200.sharded import shard_batch
    from arco_b200.synth import CaseSpec, exact_case, make_bank
    spec = CaseSpec("dist", 2, 2, 4, (32, 32), 16, queries=32, negatives=8, bank_init="fill:60",
                    caps=[80, 70, 70, 70], label_mode="absent:3" if scenario == 0 else "iid", seed=31 + scenario)
    x = exact_case(spec, 0)
    mine = shard_batch(x, spec.n_lab, rank, world)
    if scenario == 1 and rank == 0:
        # class 1 lives on rank 1 only: rank 0's local valid list [0, 2, 3] differs from the global [0, 1, 2, 3], so its
        # speculative rank-local sampler run must be redone after arco_replan_global (plan.replanned)
        for key in ("label_l", "label_u"):
            lab = mine[key].clone()
            lab[:, 0] = lab[:, 0] | lab[:, 1]
            lab[:, 1] = 0
            mine[key] = lab
    bank_g, ptr_g, caps = make_bank(spec)
    bank_c, ptr_c, _ = make_bank(spec)
    g = {k: v.to(dev) for k, v in mine.items()}
    rep_g = g["rep"].clone().requires_grad_(True)
    dbg = {}
    nk, loss = arco_b200.compute_contra_memobank_loss(
        rep_g, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"], g["high_mask"],
        bank_g, ptr_g, caps, g["rep_teacher"], delta_n=0.97, func="smc", num_queries=spec.queries,
        num_negatives=spec.negatives, process_group=dist.group.WORLD, seed=5, _debug=dbg)
    loss.backward()
    torch.cuda.synchronize()
    arco_b200.synchronize_bank(bank_g)
    plan = bank_g[0].bank.last_plan
    active = [j for j in range(spec.classes) if plan.slot_active[j]]
    replay = []
    for j in active:
        replay += [dbg["idx_anchor"][j].long().cpu(), dbg["idx_neg"][j, : spec.queries * spec.negatives].long().cpu()]
    it = iter(replay)
    glob = dbg["proto_sums"].cpu()                        # all-reduced on the device

    rep_c = mine["rep"].clone().requires_grad_(True)
    res = oracle.contra_memobank_loss(
        rep_c, mine["label_l"], mine["label_u"], mine["prob_l"], mine["prob_u"], mine["low_mask"], mine["high_mask"],
        bank_c, ptr_c, caps, mine["rep_teacher"], delta_n=0.97, sampler=lambda h, s: next(it),
        num_queries=spec.queries, num_negatives=spec.negatives, proto_sum_hook=lambda local: glob.clone())
    res.loss.backward()
    # the production path (one arco_forward call with the peer-memory exchange inside) must reproduce the staged one
    bank_f, ptr_f, _ = make_bank(spec)
    rep_f = g["rep"].clone().requires_grad_(True)
    nk_f, loss_f = arco_b200.compute_contra_memobank_loss(
        rep_f, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"], g["high_mask"],
        bank_f, ptr_f, caps, g["rep_teacher"], delta_n=0.97, func="smc", num_queries=spec.queries,
        num_negatives=spec.negatives, process_group=dist.group.WORLD, seed=5)
    loss_f.backward()
    fused = dict(keys=list(nk_f) == list(nk), loss=(float(loss_f.detach()), float(loss.detach())),
                 grad=float((rep_f.grad - rep_g.grad).abs().max()))
    # duplicate anchors accumulate through float atomics (order not fixed): the gradient may differ in the last bits
    fused_ok = fused["keys"] and fused["loss"][0] == fused["loss"][1] and fused["grad"] <= 1e-6 * float(rep_g.grad.abs().max())
    nv = int(plan.n_valid)
    ok = (
        fused_ok and
        [int(plan.valid_class[i]) for i in range(nv)] == res.valid_classes
        and list(nk) == res.new_keys
        and abs(float(loss.detach()) - float(res.loss.detach())) <= 1e-5 * max(1.0, abs(float(res.loss.detach())))
        and float((rep_g.grad.cpu() - rep_c.grad).norm() / rep_c.grad.norm()) <= 1e-5
    )
    return dict(ok=bool(ok), loss=float(loss.detach()), oracle=float(res.loss.detach()), valid=res.valid_classes,
                replanned=int(plan.replanned), fused=fused)


def _scenario_producers(rank, world, dev):
    """SURVEY 8(f) rank 2 on batch shards: the x-space class sums are all-reduced, then the teacher weight is applied."""
    from arco_b200 import producers
    from arco_b200.sharded import shard_batch
    from arco_b200.synth import CaseSpec, exact_case, make_bank
    from oracle import producers_oracle as po
    import arco_b200
    spec = CaseSpec("dist_prod", 2, 2, 4, (32, 32), 48, queries=32, negatives=8, bank_init="fill:60", caps=[80, 70, 70, 70], seed=77)
    x = exact_case(spec, 0)
    mine = shard_batch(x, spec.n_lab, rank, world)
    gen = torch.Generator().manual_seed(9)
    ws = [torch.randn(48, 48, generator=gen) / 7 for _ in range(4)]
    g = {k: v.to(dev) for k, v in mine.items()}
    bank_g, ptr_g, caps = make_bank(spec)
    bank_c, ptr_c, _ = make_bank(spec)
    xs = g["rep"].clone().requires_grad_(True)
    dbg = {}
    nk, loss = producers.compute_contra_memobank_loss_from_features(
        xs, g["rep_teacher"], [w.to(dev) for w in ws[:3]], ws[3].to(dev), g["label_l"], g["label_u"], g["prob_l"], g["prob_u"],
        g["low_mask"], g["high_mask"], bank_g, ptr_g, caps, delta_n=0.97, func="smc", num_queries=spec.queries,
        num_negatives=spec.negatives, process_group=dist.group.WORLD, seed=5, _debug=dbg)
    loss.backward()
    torch.cuda.synchronize()
    arco_b200.synchronize_bank(bank_g)
    plan = bank_g[0].bank.last_plan
    replay = []
    for j in range(spec.classes):
        if plan.slot_active[j]:
            replay += [dbg["idx_anchor"][j].long().cpu(), dbg["idx_neg"][j, : spec.queries * spec.negatives].long().cpu()]
    it = iter(replay)

    def global_sums(local):
        t = local.to(dev)
        dist.all_reduce(t)
        return t.cpu()

    xo = mine["rep"].clone().requires_grad_(True)
    res = po.contra_from_features(xo, mine["rep_teacher"], ws[:3], ws[3], mine["label_l"], mine["label_u"], mine["prob_l"], mine["prob_u"],
                                  mine["low_mask"], mine["high_mask"], bank_c, ptr_c, caps, delta_n=0.97, sampler=lambda h, s: next(it),
                                  num_queries=spec.queries, num_negatives=spec.negatives, proto_sum_hook=global_sums)
    res.loss.backward()
    lo = float(res.loss.detach())
    ok = (list(nk) == res.new_keys and abs(float(loss.detach()) - lo) <= 1e-5 * max(1.0, abs(lo))
          and float((xs.grad.cpu() - xo.grad).norm() / xo.grad.norm()) <= 1e-5)
    return dict(ok=bool(ok), loss=float(loss.detach()), oracle=lo)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        out[rank] = [_scenario(sc, rank, world, dev) for sc in range(2)] + [_scenario_producers(rank, world, dev)]
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_loss_matches_oracle():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/dist_res.txt", "w") as f:
            f.write(repr(res))
    for sc in range(2):
        assert all(res[r][sc]["ok"] for r in range(world)), res
        assert res[0][sc]["valid"] == res[1][sc]["valid"]
    assert all(res[r][2]["ok"] for r in range(world)), res                       # producers folded in, x-space sums all-reduced
    assert res[0][0]["replanned"] == 0 and res[1][0]["replanned"] == 0          # same valid list everywhere
    assert res[0][1]["replanned"] == 1 and res[0][1]["valid"] == [0, 1, 2, 3]    # rank 0 had to redraw
