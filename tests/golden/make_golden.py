#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported unmodified from ``/root/reference/code``; the only shim is
harness-side: ``torch.Tensor.cuda`` becomes the identity so the hard-coded ``.cuda()`` calls
(loss_helper_3d.py:427,433,456,466,485,508) run on a CPU-only box.  The sampler symbols are
wrapped to RECORD the indices they return (they are looked up from module globals at call time,
loss_helper_3d.py:327-334), so a parity test can replay exactly the same indices.

Outputs: one ``<case>.npz`` per entry of ``tests/cases.py:CASES`` plus ``samplers.npz``.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/code")

torch.Tensor.cuda = lambda self, *a, **k: self      # harness-side shim, see docstring

import loss_helper as ref3d        # noqa: E402  5-D volumes  (file names are inverted)
import loss_helper_3d as ref2d     # noqa: E402  4-D images

from arco_b200.synth import exact_case, make_bank   # noqa: E402
from cases import CASES, PREPARE_CASES, SAMPLER_CASES, prepare_inputs   # noqa: E402


def run_case(spec):
    mod = ref3d if len(spec.spatial) == 3 else ref2d
    calls = []

    def recording(fn):
        def wrapped(high, shape, *a, **k):
            out = fn(high, shape, *a, **k)
            n = shape[0] if isinstance(shape, tuple) else shape
            calls.append((int(high), int(n), out.clone().numpy().astype(np.int64)))
            return out
        return wrapped

    originals = (mod.grid_monte_carlo_sample, mod.grid_as_monte_carlo_sample, torch.randint)
    memobank, ptrs, caps = make_bank(spec)
    out = {}
    torch.manual_seed(spec.seed)
    random.seed(spec.seed)
    np.random.seed(spec.seed)
    try:
        for step in range(spec.steps):
            x = exact_case(spec, step)
            rep = x["rep"].clone().requires_grad_(True)
            calls.clear()
            if spec.func in ("smc", "asmc"):
                mod.grid_monte_carlo_sample = recording(originals[0])
                mod.grid_as_monte_carlo_sample = recording(originals[1])
            else:
                mod.torch.randint = recording(originals[2])
            try:
                ret = mod.compute_contra_memobank_loss(
                    rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
                    memobank, ptrs, caps, x["rep_teacher"], momentum_prototype=x.get("momentum_prototype"),
                    i_iter=spec.i_iter, delta_n=spec.delta_n, func=spec.func,
                    num_queries=spec.queries, num_negatives=spec.negatives, temp=spec.temp)
                new_keys, loss = ret[-2], ret[-1]
            finally:
                mod.grid_monte_carlo_sample, mod.grid_as_monte_carlo_sample = originals[0], originals[1]
                mod.torch.randint = originals[2]
            loss.backward()
            p = f"s{step}_"
            out[p + "new_keys"] = np.asarray(new_keys, np.int64)
            out[p + "loss"] = loss.detach().float().numpy()
            if len(ret) == 3:
                out[p + "prototype"] = ret[0].detach().float().numpy()
            out[p + "grad"] = rep.grad.float().numpy()
            out[p + "ptr"] = np.asarray([int(q) for q in ptrs], np.int64)
            out[p + "bank_len"] = np.asarray([m[0].shape[0] for m in memobank], np.int64)
            for c, m in enumerate(memobank):
                out[p + f"bank{c}"] = m[0].float().numpy()
            out[p + "call_high"] = np.asarray([c[0] for c in calls], np.int64)
            out[p + "call_shape"] = np.asarray([c[1] for c in calls], np.int64)
            for i, c in enumerate(calls):
                out[p + f"call{i}"] = c[2]
    finally:
        pass
    return out


def run_samplers():
    out = {}
    for func, high, shape, seed in SAMPLER_CASES:
        torch.manual_seed(seed)
        random.seed(seed)
        np.random.seed(seed)
        fn = ref2d.grid_monte_carlo_sample if func == "smc" else ref2d.grid_as_monte_carlo_sample
        out[f"{func}_{high}_{shape}_{seed}"] = fn(high, shape).numpy().astype(np.int64)
    return out


def _trainer_block(path, three_d):
    """The trainer's own ``with torch.no_grad():`` mask-preparation block and its ``label_onehot`` def, as source text read
    from the reference tree at generation time (train_arco_2d.py:345-393 + :492-498 / train_arco_3d.py:315-353 + :463-469)."""
    import textwrap
    lines = open(path).read().split("\n")
    start = next(i for i, ln in enumerate(lines) if ln.strip() == "with torch.no_grad():" and "alpha_t" in "".join(lines[i - 4:i]))
    end = next(i for i in range(start, len(lines)) if lines[i].strip().startswith("reco_loss = compute_contra_memobank_loss"))
    block = textwrap.dedent("\n".join(lines[start:end]))
    d0 = next(i for i, ln in enumerate(lines) if ln.startswith("def label_onehot("))
    d1 = next(i for i in range(d0 + 1, len(lines)) if lines[i].startswith("def ") or lines[i].startswith("if __name__"))
    return "\n".join(lines[d0:d1]) + "\n" + block


def run_prepare_case(case):
    """Execute the reference trainer's own lines on the case's tensors (CPU) and record what they leave behind."""
    import argparse
    import torch.nn.functional as F
    x = prepare_inputs(case)
    three_d = len(case[4]) == 3
    src = _trainer_block("/root/reference/code/train_arco_3d.py" if three_d else "/root/reference/code/train_arco_2d.py", three_d)
    ns = dict(torch=torch, np=np, F=F, args=argparse.Namespace(num_classes=x["num_classes"], weak_threshold=0.7),
              pred_l=x["pred_l"], pred_u=x["pred_u"], pred_l_teacher=x["pred_l_teacher"], pred_u_teacher=x["pred_u_teacher"],
              pred_all=torch.cat((x["pred_l"], x["pred_u"])), train_l_label=x["train_l_label"],
              train_u_aug_label=x["train_u_aug_label"], train_u_aug_logits=torch.rand(x["train_u_aug_label"].shape),
              alpha_t=x["alpha_t"])
    exec(compile(src, "<reference trainer block>", "exec"), ns)
    return dict(low_thresh=np.float32(ns["low_thresh"]), high_thresh=np.float32(ns["high_thresh"]),
                low_mask_all=ns["low_mask_all"].numpy().astype(np.uint8), high_mask_all=ns["high_mask_all"].numpy().astype(np.uint8),
                prob_l_teacher=ns["prob_l_teacher"].numpy(), prob_u_teacher=ns["prob_u_teacher"].numpy(),
                entropy=ns["entropy"].numpy(), label_l=ns["label_l"].numpy().astype(np.uint8),
                label_u=ns["label_u"].numpy().astype(np.uint8))


def main():
    import warnings
    warnings.filterwarnings("ignore")
    for spec in CASES:
        res = run_case(spec)
        path = os.path.join(HERE, spec.name + ".npz")
        np.savez_compressed(path, **res)
        losses = [float(res[f"s{t}_loss"]) for t in range(spec.steps)]
        print(f"{spec.name:16s} loss={losses} new_keys={res[f's{spec.steps-1}_new_keys'].tolist()} "
              f"calls={res[f's{spec.steps-1}_call_high'].tolist()} {os.path.getsize(path)/1024:.0f} KB")
    for case in PREPARE_CASES:
        res = run_prepare_case(case)
        path = os.path.join(HERE, case[0] + ".npz")
        np.savez_compressed(path, **res)
        print(f"{case[0]:16s} thresholds=({res['low_thresh']!r}, {res['high_thresh']!r}) low={int(res['low_mask_all'].sum())} "
              f"high={int(res['high_mask_all'].sum())} {os.path.getsize(path)/1024:.0f} KB")
    path = os.path.join(HERE, "samplers.npz")
    np.savez_compressed(path, **run_samplers())
    print("samplers.npz", f"{os.path.getsize(path)/1024:.0f} KB")


if __name__ == "__main__":
    main()
