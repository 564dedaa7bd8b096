"""CPU: the producers oracle (oracle/producers_oracle.py) and the FeatureExtractor twin against golden vectors made by executing
the reference's own class / trainer statements (tests/golden/make_golden_producers.py; model_2D.py:20-55,
train_arco_2d.py:231-236, :317-333) composed with the reference loss."""
import numpy as np
import pytest
import torch

from arco_b200.synth import exact_case, make_bank
from cases import PRODUCER_CASES, producer_inputs
from oracle import producers_oracle as po
from util import Replay, load_golden


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("case", PRODUCER_CASES, ids=lambda c: c["name"])
def test_oracle_matches_reference(case):
    spec = case["spec"]
    gold = load_golden(case["name"])
    memobank, ptrs, caps = make_bank(spec)
    for step in range(spec.steps):
        x = exact_case(spec, step)
        pin = producer_inputs(case, step)
        maps_l = [t.clone().requires_grad_(True) for t in pin["maps_l"]]
        maps_u = [t.clone().requires_grad_(True) for t in pin["maps_u"]]
        w_q_fe = [w.clone().requires_grad_(True) for w in pin["w_q_fe"]]
        w_q_rep = [w.clone().requires_grad_(True) for w in pin["w_q_rep"]]
        rep, rep_t = po.representations(w_q_fe, w_q_rep, pin["w_k_fe"], maps_l, maps_u, pin["maps_l_teacher"], pin["maps_u_teacher"])
        p = f"s{step}_"
        assert _rel(rep.detach()[:, ::8, ::4, ::4], gold[p + "rep_sample"]) <= 1e-6
        assert _rel(rep_t.detach()[:, ::8, ::4, ::4], gold[p + "rep_teacher_sample"]) <= 1e-6
        replay = Replay(gold, step)
        res = po.contra_memobank_loss(rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
                                      memobank, ptrs, caps, rep_t.detach(), delta_n=spec.delta_n, sampler=replay,
                                      num_queries=spec.queries, num_negatives=spec.negatives, temp=spec.temp)
        assert replay.done()
        res.loss.backward()
        assert list(res.new_keys) == gold[p + "new_keys"].tolist()
        assert abs(float(res.loss) - float(gold[p + "loss"])) <= 1e-6 * abs(float(gold[p + "loss"]))
        for c in range(spec.classes):
            assert _rel(memobank[c][0], gold[p + f"bank{c}"]) <= 1e-6
        assert [int(t[0]) for t in ptrs] == gold[p + "ptrs"].tolist()
        assert _rel(maps_l[4].grad, gold[p + "grad_map4_l"]) <= 1e-5
        assert _rel(maps_u[4].grad, gold[p + "grad_map4_u"]) <= 1e-5
        assert _rel(maps_u[0].grad, gold[p + "grad_map0_u"]) <= 1e-5
        for name, w in (("fea4", w_q_fe[4]), ("qrep0", w_q_rep[0]), ("qrep1", w_q_rep[1]), ("fea3", w_q_fe[3])):
            assert _rel(w.grad.sum(dim=1), gold[p + f"gw_{name}_rowsum"]) <= 1e-5
            assert _rel(w.grad.sum(dim=0), gold[p + f"gw_{name}_colsum"]) <= 1e-5
            assert abs(float(w.grad.norm()) - float(gold[p + f"gw_{name}_norm"])) <= 1e-5 * float(gold[p + f"gw_{name}_norm"])


@pytest.mark.parametrize("case", PRODUCER_CASES, ids=lambda c: c["name"])
def test_feature_extractor_twin_matches_reference(case):
    """arco_b200.producers.FeatureExtractor / make_q_representation: same parameter names as the reference module (a reference
    state_dict loads) and the same forward; trunk() + fea4 == forward()."""
    from arco_b200.producers import FeatureExtractor, make_q_representation
    gold = load_golden(case["name"])
    pin = producer_inputs(case, 0)
    D = sum(case["fea_dim"])
    q_fe, k_fe, q_rep = FeatureExtractor(case["fea_dim"], D), FeatureExtractor(case["fea_dim"], D), make_q_representation(D)
    assert list(q_fe.state_dict().keys()) == [f"fea{i}.weight" for i in range(5)]
    q_fe.load_state_dict({f"fea{i}.weight": pin["w_q_fe"][i][:, :, None, None] for i in range(5)})
    k_fe.load_state_dict({f"fea{i}.weight": pin["w_k_fe"][i][:, :, None, None] for i in range(5)})
    q_rep.load_state_dict({f"{i}.weight": pin["w_q_rep"][i][:, :, None, None] for i in range(2)})
    with torch.no_grad():
        rep = torch.cat((q_rep(q_fe(pin["maps_l"])), q_rep(q_fe(pin["maps_u"]))))
        rep_t = torch.cat((k_fe(pin["maps_l_teacher"]), k_fe(pin["maps_u_teacher"])))
        assert torch.equal(q_fe.fea4(q_fe.trunk(pin["maps_u"])), q_fe(pin["maps_u"]))
    assert _rel(rep[:, ::8, ::4, ::4], gold["s0_rep_sample"]) <= 1e-6
    assert _rel(rep_t[:, ::8, ::4, ::4], gold["s0_rep_teacher_sample"]) <= 1e-6


def test_feature_extractor_3d_twin_matches_reference():
    """arco_b200.producers.FeatureExtractor_3d against the reference class's own output (model_3D.py:20-63, the 3-D trainer's
    channel plan and q_representation, train_arco_3d.py:206-213)."""
    import torch.nn as nn
    from arco_b200.producers import FeatureExtractor_3d
    from cases import producer_inputs_3d
    gold = load_golden("producers_3d")
    x = producer_inputs_3d()
    fe = FeatureExtractor_3d([128, 64, 32, 16, 16], 16)
    assert list(fe.state_dict().keys()) == [f"fea{i}.weight" for i in range(5)]
    fe.load_state_dict({f"fea{i}.weight": x["w_fe"][i][:, :, None, None, None] for i in range(5)})
    q_rep = nn.Sequential(nn.Conv3d(16, 16, kernel_size=1, bias=False), nn.Conv3d(16, 16, kernel_size=1, bias=False))
    q_rep.load_state_dict({f"{i}.weight": x["w_rep"][i][:, :, None, None, None] for i in range(2)})
    with torch.no_grad():
        fea = fe(x["maps"])
        rep = q_rep(fea)
        assert torch.equal(fe.fea4(fe.trunk(x["maps"])), fea)
    assert _rel(fea, gold["fea"]) <= 1e-6
    assert _rel(rep, gold["rep"]) <= 1e-6
