"""GPU: the logits-in op (arco_b200.compute_contra_memobank_loss_from_logits; teacher softmax, entropy masks and the label decode
inside the classify kernel) must equal -- BIT FOR BIT -- the two-call path prepare_contrast_inputs + compute_contra_memobank_loss,
whose pieces are pinned separately (tests/test_gpu_prepare.py against golden vectors made from the trainer's own lines,
tests/test_gpu_parity.py against the reference loss)."""
import ctypes as C

import numpy as np
import pytest
import torch

import arco_b200
from arco_b200.synth import CaseSpec, make_bank

pytestmark = pytest.mark.gpu

SHAPES = [
    ("acdc", 2, 2, 4, (64, 64), 64, torch.float32, 20.0),
    ("la", 1, 2, 2, (16, 16, 12), 16, torch.float32, 35.0),
    ("c8_bf16", 2, 1, 8, (32, 48), 128, torch.bfloat16, 10.0),
    ("no_unlab", 2, 0, 4, (32, 32), 32, torch.float32, 20.0),
]


def _codes(dbg):
    lay, ws = dbg["layout"], dbg["ws"]
    d = dbg["dims"]
    n = (d.n_lab + d.n_unlab) * d.space
    return ws[lay.codes: lay.codes + n].clone()


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: s[0])
def test_logits_in_equals_prepare_then_loss(shape):
    name, n_l, n_u, Cn, spatial, D, dt, alpha = shape
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(11)
    B = n_l + n_u
    rnd = lambda *sh: torch.randn(*sh, device=dev, generator=g)
    pred_l_t, pred_u_t, pred_u = rnd(n_l, Cn, *spatial) * 2, rnd(n_u, Cn, *spatial) * 2, rnd(n_u, Cn, *spatial) * 2
    lab_l = torch.randint(0, Cn, (n_l,) + spatial, device=dev, generator=g)
    lab_u = torch.randint(-1, Cn, (n_u,) + spatial, device=dev, generator=g)
    rep = rnd(B, D, *spatial).to(dt)
    rep_t = rnd(B, D, *spatial).to(dt)
    spec = CaseSpec(name, n_l, n_u, Cn, spatial, D, queries=32, negatives=16, bank_init="fill:60", caps=[80] * Cn)
    common = dict(delta_n=0.97, func="smc", num_queries=32, num_negatives=16, temp=0.5, seed=9)
    outs = []
    for path in ("two_calls", "logits"):
        bank, ptr, caps = make_bank(spec)
        r = rep.clone().requires_grad_(True)
        dbg = {}
        if path == "two_calls":
            if n_u:
                p = arco_b200.prepare_contrast_inputs(pred_u, pred_l_t, pred_u_t, lab_l, lab_u, alpha)
                ll, lu, pl, pu, lo, hi = p["label_l"], p["label_u"], p["prob_l_teacher"], p["prob_u_teacher"], p["low_mask_all"], p["high_mask_all"]
            else:
                pl, _ = arco_b200.softmax_entropy(pred_l_t)
                pu = torch.empty((0, Cn) + spatial, device=dev)
                ll, lu = lab_l, lab_u
                lo = (lab_l >= 0).float().unsqueeze(1)
                hi = lo.clone()
            keys, loss = arco_b200.compute_contra_memobank_loss(r, ll, lu, pl, pu, lo, hi, bank, ptr, caps, rep_t, _debug=dbg, **common)
        else:
            keys, loss = arco_b200.compute_contra_memobank_loss_from_logits(r, lab_l, lab_u, pred_l_t, pred_u_t, pred_u, alpha, bank, ptr,
                                                                             caps, rep_t, _debug=dbg, **common)
        loss.backward()
        torch.cuda.synchronize()
        outs.append(dict(keys=list(keys), loss=float(loss.detach()), grad=r.grad.clone(), codes=_codes(dbg),
                         rows=[bank[c][0].cpu().clone() for c in range(Cn)]))
    a, b = outs
    assert torch.equal(a["codes"], b["codes"]), "per-pixel class / flag bytes differ"
    assert a["keys"] == b["keys"]
    assert a["loss"] == b["loss"]
    assert float((a["grad"].float() - b["grad"].float()).abs().max()) <= 1e-6 * float(a["grad"].float().abs().max())   # float atomics for duplicate anchors
    for x, y in zip(a["rows"], b["rows"]):
        assert torch.equal(x, y)
    assert np.isfinite(a["loss"]) and (n_u == 0 or sum(a["keys"]) >= 0)
