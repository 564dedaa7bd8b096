"""GPU: parity AT THE PARAMETERS bench.py TIMES (VERDICT r01 "what's weak" #1).

Every workload of BASELINE.json / SURVEY.md section 8(d) at its FULL batch, the reference's Q = 256 queries and
N = 512 negatives (train_arco_2d.py:61-62), banks pre-filled to 50000 / 30000 rows as the trainers size them
(train_arco_2d.py:147-154) for two consecutive steps, so the ring wraps while it is being sampled.  The oracle's torch
ops run on the GPU (they are device agnostic) with the device sampler's own indices replayed into them; the bars are
the ones of tests/test_gpu_parity.py: integers bit-exact, fp32 loss / gradient <= 1e-5 relative (bf16 storage: 2e-2 on
the bf16-rounded gradient).

This is the configuration `infonce_mma_kernel` runs with 32 chunks per query, `proto_tc_kernel` walks 1536 tiles with ring
wrap at 50 000 rows, `proto_tc32_kernel` accumulates 4.2 M pixels of 19 classes, and fp32 partial sums cover 2.5 M low-valid
pixels.
"""
import gc

import pytest
import torch

import oracle
from arco_b200.synth import WORKLOADS, bench_bank, bench_inputs

pytestmark = pytest.mark.gpu

Q, N = 256, 512

PRODUCTION = [
    # workload, sampler (2-D trainer default smc, 3-D trainer default asmc: train_arco_2d.py:78 / train_arco_3d.py:78)
    ("acdc2d_loss", "smc"),          # config 1: 4+4, C=4, 256x256, D=64 fp32
    ("acdc2d_trainstep", "smc"),     # config 2: 12+12, C=4, 256x256, D=496 bf16 -- the bench.py headline
    ("la3d", "asmc"),                # config 3: 2+2, C=2, 112x112x80, D=16 fp32
    ("cityscapes", "smc"),           # config 4: 8+8, C=19, 512x512, D=256 fp32
    # the same two tensor-core shapes with spatially coherent entropy masks: whole 32- / 64-pixel steps of the unlabelled
    # images are skipped by the prototype pass (per-tile step masks), which must not change a single count, key or ring row
    ("acdc2d_trainstep+coherent", "smc"),
    ("cityscapes+coherent", "smc"),
]


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("workload,func", PRODUCTION, ids=[w for w, _ in PRODUCTION])
def test_benchmarked_configuration_against_oracle(workload, func):
    import arco_b200
    dev = torch.device("cuda", 0)
    workload, _, variant = workload.partition("+")
    cfg = WORKLOADS[workload]
    spec = None
    bank_g = bank_c = None
    for step in range(2):
        spec, x = bench_inputs(workload, dev, seed=101 + step, coherent=variant == "coherent")
        assert (spec.n_lab, spec.n_unlab) == (cfg["n_lab"], cfg["n_unlab"])          # the FULL batch
        if bank_g is None:
            bank_g, ptr_g, caps = bench_bank(spec, seed=3)
            bank_c, ptr_c, _ = bench_bank(spec, seed=3)
            assert caps[0] == 50000 and all(c == 30000 for c in caps[1:])
        rep_g = x["rep"].clone().requires_grad_(True)
        dbg = {}
        new_keys, loss = arco_b200.compute_contra_memobank_loss(
            rep_g, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            bank_g, ptr_g, caps, x["rep_teacher"], delta_n=0.97, func=func, num_queries=Q, num_negatives=N,
            seed=99, _debug=dbg)
        loss.backward()
        torch.cuda.synchronize()
        arco_b200.synchronize_bank(bank_g)
        plan = bank_g[0].bank.last_plan
        Cn = spec.classes
        active = [j for j in range(Cn) if plan.slot_active[j]]
        assert len(active) == Cn                                   # every class reaches the loss at these sizes
        replay = []
        for j in active:
            replay += [dbg["idx_anchor"][j].long().cpu(), dbg["idx_neg"][j, : Q * N].long().cpu()]
        pos = [0]

        def sampler(high, shape):
            idx = replay[pos[0]]
            pos[0] += 1
            assert idx.numel() == shape and int(idx.max()) < high and int(idx.min()) >= 0
            return idx

        rep_c = x["rep"].float().clone().requires_grad_(True)
        ores = oracle.contra_memobank_loss(
            rep_c, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            bank_c, ptr_c, caps, x["rep_teacher"].float(), delta_n=0.97, sampler=sampler, num_queries=Q,
            num_negatives=N, temp=0.5)
        ores.loss.backward()
        assert pos[0] == len(replay)
        # ---- integers: exact ----
        assert list(new_keys) == ores.new_keys, f"step {step}"
        assert [int(plan.lv_count[c]) for c in range(Cn)] == ores.low_valid_counts
        assert [int(plan.n_anchor[c]) for c in range(Cn)] == [len(a) for a in ores.anchor_lists]
        assert [int(plan.valid_class[i]) for i in range(int(plan.n_valid))] == ores.valid_classes
        assert [int(q) for q in ptr_g] == [int(q) for q in ptr_c]
        if Cn > 3:
            assert sum(ores.new_keys) > 0                          # the ring really wraps (banks start full)
        for c in range(Cn):
            assert torch.equal(bank_g[c][0].cpu(), bank_c[c][0].float()), f"bank {c}, step {step}"   # FIFO order, verbatim rows
        # ---- floats ----
        proto_g = (dbg["proto_sums"][:, :-1] / dbg["proto_sums"][:, -1:]).float()
        assert _rel(proto_g, ores.proto) <= 2e-5, f"step {step}"
        tol = 2e-2 if spec.dtype == "bf16" else 1e-5
        lo = float(ores.loss.detach())
        assert abs(float(loss.detach()) - lo) <= min(tol, 1e-4) * max(1.0, abs(lo)), (step, float(loss.detach()), lo)
        # logits of the first and the last active position, all Q x (1+N) of them
        for j in (active[0], active[-1]):
            assert _rel(dbg["logits"][j], ores.slots[j]["logits"]) <= 1e-5, f"logits {j}"
        assert _rel(rep_g.grad.float(), rep_c.grad) <= tol, f"step {step}"
        # gradient support == the sampled anchor pixels
        pix = dbg["anchor_pix"]
        pix = pix[pix >= 0].long().unique()
        touched = (rep_g.grad.flatten(2) != 0).any(dim=1).flatten().nonzero().flatten()
        assert torch.isin(touched, pix).all()
        del rep_g, rep_c, ores, dbg, x, replay, loss, touched, pix, proto_g
        gc.collect()
        torch.cuda.empty_cache()


MOM_CASES = [
    # workload, batch override (None = full), momentum, force prefill
    ("acdc2d_loss", None, False, False),
    ("acdc2d_loss", None, False, True),
    ("acdc2d_loss", None, True, False),
    ("acdc2d_trainstep", (2, 2), False, True),
    ("acdc2d_trainstep", (2, 2), True, False),
    ("la3d", None, False, False),
    ("cityscapes", (1, 1), False, True),
]


@pytest.mark.parametrize("workload,batch,with_momentum,prefill", MOM_CASES)
def test_fused_forward_is_bit_identical_to_staged(workload, batch, with_momentum, prefill, monkeypatch):
    """every oracle parity test drives the STAGED path (``_debug`` / ``_inject``), while bench.py and
    training drive ``_forward_fused`` -> ``arco_forward`` (C-side streams/events, one packed buffer with hand-computed
    offsets, the arco_step_io layout).  Same inputs, seed and bank state through both: loss, rep.grad, new_keys, bank
    rows, pointers and (a11) the returned prototypes must be BIT-identical, for two consecutive steps."""
    import arco_b200
    from arco_b200 import contra
    dev = torch.device("cuda", 0)
    monkeypatch.setattr(contra, "_PREFILL_GRAD", prefill)
    monkeypatch.setattr(contra, "_PREFILL_MIN_BYTES", 0)
    n_lab, n_unlab = batch if batch else (None, None)
    outs = []
    for staged in (False, True):
        res = []
        spec, x = bench_inputs(workload, dev, seed=11, n_lab=n_lab, n_unlab=n_unlab)
        bank, ptr, caps = bench_bank(spec, seed=5)
        mom = None
        if with_momentum:
            g = torch.Generator(device=dev)
            g.manual_seed(17)
            mom = torch.randn((spec.classes, Q, 1, spec.feat), device=dev, generator=g)
        for step in range(2):
            rep = x["rep"].clone().requires_grad_(True)
            kw = dict(delta_n=0.97, func="smc", num_queries=Q, num_negatives=N, seed=4242)
            if mom is not None:
                kw.update(momentum_prototype=mom, i_iter=7 + step)
            if staged:
                kw["_debug"] = {}
            out = arco_b200.compute_contra_memobank_loss(
                rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
                bank, ptr, caps, x["rep_teacher"], **kw)
            loss = out[-1]
            loss.backward()
            torch.cuda.synchronize()
            arco_b200.synchronize_bank(bank)
            res.append(dict(loss=loss.detach().clone(), grad=rep.grad.clone(), keys=list(out[-2]),
                            proto=out[0].clone() if mom is not None else None,
                            ptr=[int(q) for q in ptr], bank=[bank[c][0].clone() for c in range(spec.classes)]))
            if mom is not None:
                mom = out[0].detach()
        outs.append(res)
        del x, bank
        gc.collect()
        torch.cuda.empty_cache()
    for step, (f, s) in enumerate(zip(*outs)):
        assert torch.equal(f["loss"], s["loss"]), (step, float(f["loss"]), float(s["loss"]))
        assert f["keys"] == s["keys"] and f["ptr"] == s["ptr"], step
        assert torch.equal(f["grad"], s["grad"]), step
        for a, b in zip(f["bank"], s["bank"]):
            assert torch.equal(a, b), step
        if f["proto"] is not None:
            assert torch.equal(f["proto"], s["proto"]), step


def test_host_mirror_follows_steps_without_synchronize_bank():
    """``queue_prtlis`` and the label-error status must reach the host in a plain training loop
    that only ever indexes ``[-1]`` -- no ``synchronize_bank``, no ``new_keys`` access.  The last CTA of every step
    mirrors its plan into pinned memory; ``poll()`` at the start of the next call applies what has landed."""
    import arco_b200
    dev = torch.device("cuda", 0)
    spec, x = bench_inputs("acdc2d_loss", dev, seed=21, n_lab=2, n_unlab=2)
    caps = [700, 500, 500, 500]
    bank = [[torch.zeros(1, spec.feat)] for _ in range(spec.classes)]
    ptr = [torch.zeros(1, dtype=torch.long) for _ in range(spec.classes)]
    bank_c = [[torch.zeros(1, spec.feat)] for _ in range(spec.classes)]
    ptr_c = [torch.zeros(1, dtype=torch.long) for _ in range(spec.classes)]
    expect = []
    for step in range(4):
        rep = x["rep"].clone().requires_grad_(True)
        loss = arco_b200.compute_contra_memobank_loss(
            rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            bank, ptr, caps, x["rep_teacher"], delta_n=0.97, func="smc", num_queries=32, num_negatives=16, seed=1)[-1]
        loss.backward()
        torch.cuda.synchronize()                               # the step is finished on the device; the HOST mirror is
        bank[0].bank.poll()                                    # refreshed by the same non-blocking call the op makes
        ores = oracle.contra_memobank_loss(
            x["rep"].float(), x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            bank_c, ptr_c, caps, x["rep_teacher"].float(), delta_n=0.97,
            sampler=lambda h, s: torch.zeros(s, dtype=torch.long), num_queries=32, num_negatives=16)
        expect.append([int(q) for q in ptr_c])
        assert [int(q) for q in ptr] == expect[-1], f"queue_prtlis after step {step}"
        assert bank[0].bank.host_len == [b[0].shape[0] for b in bank_c]
    # a class id >= C in an integer label map: silently dropped pixels in round 1, now raised by the NEXT call
    bad = x["labels"][spec.n_lab:].clone()
    bad[0, 0, 0] = spec.classes + 3
    arco_b200.compute_contra_memobank_loss(
        x["rep"], x["labels"][: spec.n_lab].contiguous(), bad, x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
        bank, ptr, caps, x["rep_teacher"], delta_n=0.97, func="smc", num_queries=32, num_negatives=16, seed=1)
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="class id"):
        arco_b200.compute_contra_memobank_loss(
            x["rep"], x["labels"][: spec.n_lab].contiguous(), x["labels"][spec.n_lab:].contiguous(), x["prob_l"],
            x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps, x["rep_teacher"], delta_n=0.97, func="smc",
            num_queries=32, num_negatives=16, seed=1)


def test_out_of_range_injected_index_is_flagged():
    """an out-of-range sample index used to be clamped silently; it now raises ARCO_ST_INDEX_RANGE."""
    import arco_b200
    dev = torch.device("cuda", 0)
    spec, x = bench_inputs("acdc2d_loss", dev, seed=31, n_lab=1, n_unlab=1)
    bank, ptr, caps = bench_bank(spec, seed=5)
    q, n = 8, 4
    anchors = [torch.zeros(q, dtype=torch.long) for _ in range(spec.classes)]
    negs = [torch.zeros(q * n, dtype=torch.long) for _ in range(spec.classes)]
    negs[1][3] = 10 ** 6                                         # beyond the 30000-row ring
    keys, _ = arco_b200.compute_contra_memobank_loss(
        x["rep"], x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps,
        x["rep_teacher"], delta_n=0.97, num_queries=q, num_negatives=n, _inject={"anchor": anchors, "neg": negs})
    with pytest.raises(ValueError, match="sample index"):
        list(keys)


@pytest.mark.parametrize("workload,batch", [("acdc2d_loss", None), ("acdc2d_trainstep", (2, 2))])
def test_sparse_grad_contract_matches_the_dense_gradient(workload, batch):
    """Opt-in ``sparse_grad=True``: the op-owned gradient buffer (previous step's anchor pixels cleared, no dense zero
    fill) must hold exactly the values of the default dense ``grad_rep``, step after step."""
    import arco_b200
    dev = torch.device("cuda", 0)
    n_lab, n_unlab = batch if batch else (None, None)
    grads = {}
    for sparse in (False, True):
        spec, x = bench_inputs(workload, dev, seed=13, n_lab=n_lab, n_unlab=n_unlab)
        bank, ptr, caps = bench_bank(spec, seed=5)
        got = []
        for step in range(3):
            leaf = x["rep"].clone().requires_grad_(True)
            rep = leaf * 1                                       # non-leaf, as in the trainers (output of q_representation)
            seen = []
            rep.register_hook(lambda g, seen=seen: seen.append(g.clone()))
            _, loss = arco_b200.compute_contra_memobank_loss(
                rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps,
                x["rep_teacher"], delta_n=0.97, func="smc", num_queries=Q, num_negatives=N, seed=4242 + step,
                sparse_grad=sparse)
            loss.backward()
            torch.cuda.synchronize()
            got.append((loss.detach().clone(), seen[0]))
        grads[sparse] = got
        del x, bank
        gc.collect()
        torch.cuda.empty_cache()
    for (l0, g0), (l1, g1) in zip(grads[False], grads[True]):
        assert torch.equal(l0, l1)
        assert torch.equal(g0, g1)
        assert float(g0.float().abs().sum()) > 0


@pytest.mark.parametrize("workload,batch", [("acdc2d_loss", None), ("cityscapes", (1, 1)), ("la3d", (1, 1))])
def test_one_launch_classify_plan_equals_the_two_launch_form(workload, batch):
    """arco_classify_plan (scan + plan in the classify kernel's tail, self-cleaning counters, warp-parallel plan) against
    the stand-alone arco_classify_count + arco_scan_plan pair: identical plan words, tile offsets, bank bookkeeping and
    loss, for three consecutive steps on the same bank (the counters must come back to zero every step)."""
    import ctypes as C

    import arco_b200
    from arco_b200 import _cabi
    dev = torch.device("cuda", 0)
    n_lab, n_unlab = batch if batch else (None, None)
    outs = []
    for legacy in (False, True):
        spec, x = bench_inputs(workload, dev, seed=19, n_lab=n_lab, n_unlab=n_unlab)
        bank, ptr, caps = bench_bank(spec, seed=5)
        res = []
        for step in range(3):
            dbg = {"legacy_scan": legacy}
            _, loss = arco_b200.compute_contra_memobank_loss(
                x["rep"], x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps,
                x["rep_teacher"], delta_n=0.97, func="smc", num_queries=64, num_negatives=32, seed=7, _debug=dbg)
            torch.cuda.synchronize()
            L = dbg["layout"]
            ws = dbg["ws"]
            plan = _cabi.Plan.from_buffer_copy(ws[L.plan: L.plan + C.sizeof(_cabi.Plan)].cpu().numpy().tobytes())
            fields = {}
            for name, _t in _cabi.Plan._fields_:
                if name in ("scan_done", "loss_done", "proto_done", "proto_done2", "reserved"):
                    continue
                v = getattr(plan, name)
                fields[name] = list(v) if hasattr(v, "__len__") else v
            n_off = spec.classes * (L.n_tiles + 1) * 4
            res.append(dict(plan=fields, loss=loss.detach().clone(),
                            off_a=ws[L.off_anchor: L.off_anchor + n_off].clone(), off_k=ws[L.off_key: L.off_key + n_off].clone(),
                            counters=bank[0].bank._counters.clone(), ptr=[int(q) for q in ptr]))
        outs.append(res)
        del x, bank
        gc.collect()
        torch.cuda.empty_cache()
    for step, (a, b) in enumerate(zip(*outs)):
        assert a["plan"] == b["plan"], step
        assert torch.equal(a["off_a"], b["off_a"]) and torch.equal(a["off_k"], b["off_k"]), step
        assert torch.equal(a["loss"], b["loss"]) and a["ptr"] == b["ptr"], step
        ctr = a["counters"].clone()
        assert int(ctr[40]) == step + 1                            # ARCO_CTR_STEP: the device step counter
        ctr[40] = 0
        assert int(ctr.abs().sum()) == 0, "arco_classify_plan must leave its tickets / accumulators zero"


def test_whole_step_replays_as_a_cuda_graph():
    """Nothing in the launch parameters changes from step to step (Philox stream and host-mirror slot come from the
    bank's device step counter), so forward + backward can be captured once with ``torch.cuda.graph`` and replayed:
    every replay must equal the eager step with the same bank history -- loss, gradient, bank rows -- and the host
    mirror must keep following (queue pointers, sequence numbers)."""
    import arco_b200
    dev = torch.device("cuda", 0)
    runs = {}
    for mode in ("eager", "graph"):
        spec, x = bench_inputs("acdc2d_loss", dev, seed=23, n_lab=2, n_unlab=2)
        bank, ptr, caps = bench_bank(spec, seed=5)
        rep = x["rep"].clone().requires_grad_(True)

        def step():
            rep.grad = None
            _, loss = arco_b200.compute_contra_memobank_loss(
                rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps,
                x["rep_teacher"], delta_n=0.97, func="smc", num_queries=64, num_negatives=32, seed=77)
            loss.backward()
            return loss

        out = []
        step()                                                     # step 1 eager in both modes (adoption, lazy init)
        torch.cuda.synchronize()
        if mode == "eager":
            for _ in range(3):
                loss = step()
                torch.cuda.synchronize()
                out.append((loss.detach().clone(), rep.grad.clone(), [int(q) for q in ptr]))
        else:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    loss = step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for _ in range(3):
                g.replay()
                torch.cuda.synchronize()
                bank[0].bank.poll()
                out.append((loss.detach().clone(), rep.grad.clone(), [int(q) for q in ptr]))
            assert bank[0].bank._applied == 4                      # device sequence numbers 1..4 landed on the host
        arco_b200.synchronize_bank(bank)
        runs[mode] = (out, [bank[c][0].clone() for c in range(spec.classes)])
    for (l0, g0, p0), (l1, g1, p1) in zip(runs["eager"][0], runs["graph"][0]):
        assert torch.equal(l0, l1) and torch.equal(g0, g1) and p0 == p1
    for a, b in zip(runs["eager"][1], runs["graph"][1]):
        assert torch.equal(a, b)
    assert not torch.equal(runs["graph"][0][0][1], runs["graph"][0][1][1])     # replays draw fresh sample streams
