"""GPU: the replay cache of arco_forward (forward.cu) changes nothing but the host cost of a step.

Two identical banks run the same sequence of steps through the public op, one with the cache on and one with it off.  The
input TENSORS stay the same objects (same addresses, as a training loop's allocator hands them back) while their CONTENTS
change every step, so a replayed graph must read this step's data through last step's pointers.  Everything the step
returns or mutates must be bit-identical between the two runs: loss, rep.grad, new_keys, ring rows, queue pointers."""
import ctypes as C

import pytest
import torch

from arco_b200.synth import CaseSpec, exact_case, make_bank

pytestmark = pytest.mark.gpu

SPECS = [
    CaseSpec("replay_pipe_f32", 2, 2, 4, (48, 48), 64, queries=32, negatives=16, bank_init="fill:100", caps=[150] * 4),
    CaseSpec("replay_tc_bf16", 1, 2, 5, (32, 32), 128, queries=8, negatives=8, dtype="bf16", bank_init="fill:10",
             caps=[16, 12, 12, 12, 12], mask_frac=0.9),
    CaseSpec("replay_small_la", 1, 1, 2, (16, 16, 12), 16, queries=16, negatives=8, bank_init="randn1", func="asmc"),
]


def _stats():
    from arco_b200 import _cabi
    buf = (C.c_int64 * 3)()
    _cabi.check(_cabi.lib.arco_forward_replay_stats(buf), "arco_forward_replay_stats")
    return list(buf)


def _run(spec, replay_on, seed, steps=10):
    import arco_b200
    from arco_b200 import _cabi
    dev = torch.device("cuda", 0)
    prev = _cabi.lib.arco_forward_replay(1 if replay_on else 0)
    try:
        bank, ptr, caps = make_bank(spec)
        if spec.dtype == "bf16":
            for m in bank:
                m[0] = m[0].to(torch.bfloat16).to(torch.float32)
        x0 = {k: v.to(dev) for k, v in exact_case(spec, 0).items()}
        rep = x0["rep"].clone().requires_grad_(True)
        before = _stats()
        out = []
        for step in range(steps):
            x = exact_case(spec, step)
            with torch.no_grad():
                for k, v in x.items():                       # new contents, same tensors
                    if k == "rep":
                        rep.copy_(v)
                    else:
                        x0[k].copy_(v)
            rep.grad = None
            new_keys, loss = arco_b200.compute_contra_memobank_loss(
                rep, x0["label_l"], x0["label_u"], x0["prob_l"], x0["prob_u"], x0["low_mask"], x0["high_mask"], bank, ptr,
                caps, x0["rep_teacher"], delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries,
                num_negatives=spec.negatives, temp=spec.temp, seed=seed)
            loss.backward()
            torch.cuda.synchronize()
            arco_b200.synchronize_bank(bank)
            out.append(dict(loss=loss.detach().clone(), grad=rep.grad.clone(), new_keys=[int(k) for k in new_keys],
                            ptr=[int(q) for q in ptr], rows=[bank[c][0].cpu().clone() for c in range(spec.classes)]))
        after = _stats()
        return out, [a - b for a, b in zip(after, before)]
    finally:
        _cabi.lib.arco_forward_replay(prev)


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: s.name)
@pytest.mark.parametrize("seed", [123, None], ids=["seeded", "default_rng"])
def test_replayed_steps_equal_direct_launches(spec, seed):
    if seed is None:
        torch.cuda.manual_seed(4242)
    on, d_on = _run(spec, True, seed)
    if seed is None:
        torch.cuda.manual_seed(4242)
        # the default-seed path keys its Philox stream by the order in which banks first met it
        import arco_b200.contra as contra
        contra._BANK_SERIAL -= 1
    off, d_off = _run(spec, False, seed)
    assert d_off[0] == 0 and d_off[1] == 0                       # cache off: nothing captured, nothing replayed
    assert d_on[0] >= 2, d_on                                    # cache on: several replayed steps (the graph may stem from an earlier test)
    for a, b in zip(on, off):
        assert torch.equal(a["loss"], b["loss"])
        assert torch.equal(a["grad"], b["grad"])
        assert a["new_keys"] == b["new_keys"] and a["ptr"] == b["ptr"]
        for ra, rb in zip(a["rows"], b["rows"]):
            assert torch.equal(ra, rb)


def test_capturing_stream_runs_the_launches_directly():
    """Inside torch.cuda.graph the caller's stream is already capturing: arco_forward must not start a capture of its own."""
    import arco_b200
    from arco_b200 import _cabi
    dev = torch.device("cuda", 0)
    spec = SPECS[0]
    bank, ptr, caps = make_bank(spec)
    x = {k: v.to(dev) for k, v in exact_case(spec, 0).items()}
    rep = x["rep"].clone().requires_grad_(True)

    def step():
        rep.grad = None
        _, loss = arco_b200.compute_contra_memobank_loss(
            rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps,
            x["rep_teacher"], delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries, num_negatives=spec.negatives,
            temp=spec.temp, seed=5)
        loss.backward()
        return loss

    prev = _cabi.lib.arco_forward_replay(1)
    try:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        before = _stats()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                loss = step()
        torch.cuda.current_stream(dev).wait_stream(side)
        assert _stats() == before                                # neither replayed, nor captured, nor counted as a sighting
        g.replay()
        torch.cuda.synchronize()
        assert torch.isfinite(loss).all()
    finally:
        _cabi.lib.arco_forward_replay(prev)
