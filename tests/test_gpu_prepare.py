"""GPU: device-side mask / threshold preparation (arco_softmax_rows, arco_entropy_masks -- SURVEY.md section 8(f) rank 1)
against the golden vectors made from the reference trainers' own lines and against numpy/torch at larger sizes.
Thresholds and masks are bit-exact given the same entropies; probabilities / entropies agree to fp32 libm rounding."""
import os

import numpy as np
import pytest
import torch

from cases import PREPARE_CASES, prepare_inputs
from oracle import prepare_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", PREPARE_CASES, ids=lambda c: c[0])
def test_masks_and_thresholds_bit_exact_on_reference_entropy(case):
    from arco_b200.prepare import entropy_masks
    dev = torch.device("cuda", 0)
    gold = np.load(os.path.join(GOLD, case[0] + ".npz"))
    x = prepare_inputs(case)
    low, high, thr = entropy_masks(torch.from_numpy(gold["entropy"]).to(dev), x["train_l_label"].to(dev),
                                   x["train_u_aug_label"].to(dev), x["alpha_t"])
    thr = thr.cpu().numpy()
    assert thr[0].tobytes() == gold["low_thresh"].tobytes() and thr[1].tobytes() == gold["high_thresh"].tobytes()
    assert np.array_equal(low.cpu().numpy().astype(np.uint8), gold["low_mask_all"])
    assert np.array_equal(high.cpu().numpy().astype(np.uint8), gold["high_mask_all"])


@pytest.mark.parametrize("case", PREPARE_CASES, ids=lambda c: c[0])
def test_prepare_end_to_end(case):
    import arco_b200
    dev = torch.device("cuda", 0)
    gold = np.load(os.path.join(GOLD, case[0] + ".npz"))
    x = prepare_inputs(case)
    out = arco_b200.prepare_contrast_inputs(x["pred_u"].to(dev), x["pred_l_teacher"].to(dev), x["pred_u_teacher"].to(dev),
                                            x["train_l_label"].to(dev), x["train_u_aug_label"].to(dev), x["alpha_t"])
    assert np.abs(out["prob_l_teacher"].cpu().numpy() - gold["prob_l_teacher"]).max() <= 1e-6
    assert np.abs(out["prob_u_teacher"].cpu().numpy() - gold["prob_u_teacher"]).max() <= 1e-6
    ent = out["entropy"].cpu().numpy()
    assert np.abs(ent - gold["entropy"]).max() <= 2e-6
    thr = out["thresholds"].cpu().numpy()
    assert abs(float(thr[0]) - float(gold["low_thresh"])) <= 2e-6 and abs(float(thr[1]) - float(gold["high_thresh"])) <= 2e-6
    # a pixel may change side only if its entropy sits within rounding distance of the threshold
    n_lab = x["train_l_label"].shape[0]
    for key, t in (("low_mask_all", gold["low_thresh"]), ("high_mask_all", gold["high_thresh"])):
        got = out[key].cpu().numpy().astype(np.uint8)
        diff = got != gold[key]
        assert not diff[:n_lab].any()
        assert diff.sum() <= 2
        assert np.all(np.abs(gold["entropy"][diff[n_lab:, 0]] - t) <= 4e-6)


@pytest.mark.parametrize("n,alpha,ignore", [(1_500_000, 14.0, 0.05), (786_432, 20.0, 0.0), (300_001, 0.37, 0.5),
                                            (4097, 0.0, 0.0), (1, 12.0, 0.0), (50_000, 20.0, 1.0)])
def test_radix_select_percentiles_match_numpy(n, alpha, ignore):
    from arco_b200.prepare import entropy_masks
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(n)
    ent = (torch.rand(1, n, generator=g) * 1.4 - 0.01)
    if n > 1000:
        ent[0, ::7] = ent[0, 3]                                          # heavy ties
    lab_u = torch.randint(0, 4, (1, n), generator=g)
    lab_u[torch.rand(1, n, generator=g) < ignore] = -1
    lab_l = torch.randint(-1, 4, (2, n), generator=g)
    low, high, thr = entropy_masks(ent.to(dev), lab_l.to(dev), lab_u.to(dev), alpha)
    if int((lab_u >= 0).sum()) == 0:
        assert bool(torch.isnan(thr).all()) and float(low[2:].sum()) == 0 and float(high[2:].sum()) == 0
        return
    rl, rh, lt, ht = prepare_oracle.masks_from_entropy(ent, lab_l, lab_u, alpha)
    thr = thr.cpu().numpy()
    assert thr[0].tobytes() == np.float32(lt).tobytes() and thr[1].tobytes() == np.float32(ht).tobytes()
    assert torch.equal(low.cpu(), rl) and torch.equal(high.cpu(), rh)


def test_prepared_inputs_feed_the_loss():
    """The prepared tensors (integer label maps, device masks) drive the loss op to the same value as the one-hot /
    mask tensors the reference block builds (same sampler seed)."""
    import arco_b200
    from arco_b200.synth import CaseSpec, exact_case, make_bank
    dev = torch.device("cuda", 0)
    case = PREPARE_CASES[0]
    x = prepare_inputs(case)
    gold = np.load(os.path.join(GOLD, case[0] + ".npz"))
    spec = CaseSpec("prep_loss", case[1], case[2], case[3], case[4], 16, queries=16, negatives=8, bank_init="fill:40", caps=[60] * 4, seed=9)
    reps = exact_case(spec, 0)
    out = arco_b200.prepare_contrast_inputs(x["pred_u"].to(dev), x["pred_l_teacher"].to(dev), x["pred_u_teacher"].to(dev),
                                            x["train_l_label"].to(dev), x["train_u_aug_label"].to(dev), x["alpha_t"])
    losses = []
    for variant in range(2):
        bank, ptr, caps = make_bank(spec)
        rep = reps["rep"].to(dev).clone().requires_grad_(True)
        if variant == 0:
            args = (out["label_l"], out["label_u"], out["prob_l_teacher"], out["prob_u_teacher"], out["low_mask_all"], out["high_mask_all"])
        else:
            args = (torch.from_numpy(gold["label_l"]).long().to(dev), torch.from_numpy(gold["label_u"]).long().to(dev),
                    torch.from_numpy(gold["prob_l_teacher"]).to(dev), torch.from_numpy(gold["prob_u_teacher"]).to(dev),
                    torch.from_numpy(gold["low_mask_all"]).float().to(dev), torch.from_numpy(gold["high_mask_all"]).float().to(dev))
        nk, loss = arco_b200.compute_contra_memobank_loss(rep, *args, bank, ptr, caps, reps["rep_teacher"].to(dev), delta_n=0.97,
                                                          func="smc", num_queries=16, num_negatives=8, seed=5)
        losses.append((list(nk), float(loss.detach())))
    assert losses[0][0] == losses[1][0]
    assert abs(losses[0][1] - losses[1][1]) <= 1e-5 * max(1.0, abs(losses[1][1]))
