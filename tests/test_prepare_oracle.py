"""CPU: the oracle's restatement of the trainers' mask / threshold preparation (oracle/prepare_oracle.py) against golden
vectors made by EXECUTING the reference's own lines (tests/golden/make_golden.py: train_arco_2d.py:345-393 + :492-498,
train_arco_3d.py:315-353 + :463-469), and the float32 restatement of np.percentile the CUDA kernel implements."""
import os

import numpy as np
import pytest
import torch

from cases import PREPARE_CASES, prepare_inputs
from oracle import prepare_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", PREPARE_CASES, ids=lambda c: c[0])
def test_prepare_oracle_matches_reference_lines(case):
    gold = np.load(os.path.join(GOLD, case[0] + ".npz"))
    x = prepare_inputs(case)
    out = prepare_oracle.prepare(x["pred_u"], x["pred_l_teacher"], x["pred_u_teacher"], x["train_l_label"],
                                 x["train_u_aug_label"], x["alpha_t"], x["num_classes"])
    assert np.array_equal(out["label_l"].numpy().astype(np.uint8), gold["label_l"])
    assert np.array_equal(out["label_u"].numpy().astype(np.uint8), gold["label_u"])
    # libm / SIMD width may differ between hosts: probabilities to 1e-6, the threshold logic exactly on the golden entropy
    assert np.abs(out["prob_l_teacher"].numpy() - gold["prob_l_teacher"]).max() <= 1e-6
    assert np.abs(out["prob_u_teacher"].numpy() - gold["prob_u_teacher"]).max() <= 1e-6
    assert np.abs(out["entropy"].numpy() - gold["entropy"]).max() <= 1e-6
    low, high, lt, ht = prepare_oracle.masks_from_entropy(torch.from_numpy(gold["entropy"]), x["train_l_label"],
                                                          x["train_u_aug_label"], x["alpha_t"])
    assert np.float32(lt).tobytes() == gold["low_thresh"].tobytes() and np.float32(ht).tobytes() == gold["high_thresh"].tobytes()
    assert np.array_equal(low.numpy().astype(np.uint8), gold["low_mask_all"])
    assert np.array_equal(high.numpy().astype(np.uint8), gold["high_mask_all"])


def percentile_f32(sorted_vals: np.ndarray, percent: float) -> np.float32:
    """The float32 arithmetic of numpy 2.x np.percentile(method='linear') that csrc/prepare.cu implements."""
    f32 = np.float32
    n = len(sorted_vals)
    q32 = f32(f32(percent) / f32(100))
    v = f32(f32(n - 1) * q32)
    fl = np.floor(v)
    g = f32(np.float64(v) - np.float64(fl))
    if v >= n - 1:
        a = b = sorted_vals[-1]
    elif v < 0:
        a = b = sorted_vals[0]
    else:
        a, b = sorted_vals[int(fl)], sorted_vals[int(fl) + 1]
    d = f32(b - a)
    return f32(b - f32(d * f32(f32(1) - g))) if g >= 0.5 else f32(a + f32(d * g))


def test_float32_percentile_restatement_is_numpy():
    rs = np.random.RandomState(7)
    for trial in range(1500):
        n = int(rs.randint(1, 2_000_000)) if trial % 100 == 0 else int(rs.randint(1, 4000))
        a = (rs.rand(n) * 3).astype(np.float32)
        if trial % 5 == 0:
            a = np.round(a * 8) / np.float32(8)                         # ties
        q = float(rs.rand() * 100) if trial % 3 else float(20 * (1 - rs.randint(0, 101) / 100))
        if trial % 4 == 0:
            q = 100 - q
        want = np.float32(np.percentile(a, q))
        assert percentile_f32(np.sort(a), q).tobytes() == want.tobytes(), (n, q)
