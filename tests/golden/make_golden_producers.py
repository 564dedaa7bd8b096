#!/usr/bin/env python
"""Golden vectors for SURVEY.md section 8(f) rank 2 (representation producers), made by executing the REFERENCE's own source:

* ``class FeatureExtractor`` is cut out of ``/root/reference/code/model_2D.py`` as text and exec'ed (the module itself cannot
  be imported here: ``networks/`` pulls in ``efficientnet_pytorch``, which is not installed);
* the statements that build ``q_representation`` and the two extractors are cut out of ``train_arco_2d.py:231-236``;
* the composition ``train_arco_2d.py:313-333`` (extractors -> q_representation -> cat) is executed from the trainer's own lines;
* the loss is the reference's ``loss_helper_3d.compute_contra_memobank_loss``, imported unmodified, with its sampler calls
  recorded for replay.

Build container only:    python tests/golden/make_golden_producers.py       -> tests/golden/producers_*.npz
"""
import os
import random
import sys
import textwrap

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/code")
torch.Tensor.cuda = lambda self, *a, **k: self      # harness-side shim (hard-coded .cuda() calls)
nn.Module.cuda = lambda self, *a, **k: self

import loss_helper_3d as ref2d                       # noqa: E402
from arco_b200.synth import exact_case, make_bank   # noqa: E402
from cases import PRODUCER_CASES, producer_inputs, producer_inputs_3d   # noqa: E402

MODEL = "/root/reference/code/model_2D.py"
TRAINER = "/root/reference/code/train_arco_2d.py"


def _class_source():
    lines = open(MODEL).read().split("\n")
    d0 = next(i for i, ln in enumerate(lines) if ln.startswith("class FeatureExtractor("))
    d1 = next(i for i in range(d0 + 1, len(lines)) if lines[i] and not lines[i][0].isspace())
    return "\n".join(lines[d0:d1])


def _trainer_lines(first_prefix, last_prefix):
    lines = open(TRAINER).read().split("\n")
    a = next(i for i, ln in enumerate(lines) if ln.strip().startswith(first_prefix))
    b = next(i for i in range(a, len(lines)) if lines[i].strip().startswith(last_prefix))
    return textwrap.dedent("\n".join(lines[a:b + 1]))


def run(case):
    spec = case["spec"]
    ns = dict(torch=torch, nn=nn, F=F, np=np)
    exec(compile(_class_source(), "<model_2D.FeatureExtractor>", "exec"), ns)
    # train_arco_2d.py:231-236 (q_representation, k_feature_extractor, q_feature_extractor), verbatim
    build = _trainer_lines("q_representation = nn.Sequential(", "q_feature_extractor = FeatureExtractor(")
    assert list(case["fea_dim"]) == [256, 128, 64, 32, 16], "the trainer's statements hard-code this channel plan"
    exec(compile(build, "<train_arco_2d.py:231-236>", "exec"), ns)
    q_rep, q_fe, k_fe = ns["q_representation"], ns["q_feature_extractor"], ns["k_feature_extractor"]
    out = {}
    memobank, ptrs, caps = make_bank(spec)
    originals = (ref2d.grid_monte_carlo_sample, ref2d.grid_as_monte_carlo_sample)
    # train_arco_2d.py:317-333 minus the dead l_feature_map_2 lines, verbatim
    compose = _trainer_lines("l_feature_all = q_feature_extractor(l_feature_map)", "pred_all_teacher = torch.cat((rep_l_teacher, rep_u_teacher))")
    compose = "\n".join(ln for ln in compose.split("\n") if "l_feature_map_2" not in ln and "pred_all = " not in ln)
    torch.manual_seed(spec.seed)
    random.seed(spec.seed)
    np.random.seed(spec.seed)
    for step in range(spec.steps):
        x = exact_case(spec, step)
        pin = producer_inputs(case, step)
        with torch.no_grad():
            for i in range(5):
                getattr(q_fe, f"fea{i}").weight.copy_(pin["w_q_fe"][i].view_as(getattr(q_fe, f"fea{i}").weight))
                getattr(k_fe, f"fea{i}").weight.copy_(pin["w_k_fe"][i].view_as(getattr(k_fe, f"fea{i}").weight))
            q_rep[0].weight.copy_(pin["w_q_rep"][0].view_as(q_rep[0].weight))
            q_rep[1].weight.copy_(pin["w_q_rep"][1].view_as(q_rep[1].weight))
        for m in (q_fe, k_fe, q_rep):
            m.zero_grad()
        maps_l = [t.clone().requires_grad_(True) for t in pin["maps_l"]]
        maps_u = [t.clone().requires_grad_(True) for t in pin["maps_u"]]
        env = dict(ns, l_feature_map=maps_l, u_feature_map=maps_u, l_feature_map_teacher=pin["maps_l_teacher"],
                   u_feature_map_teacher=pin["maps_u_teacher"])
        exec(compile(compose, "<train_arco_2d.py:317-333>", "exec"), env)
        rep_all, rep_teacher = env["rep_all"], env["pred_all_teacher"]
        calls = []

        def recording(fn):
            def wrapped(high, shape, *a, **k):
                res = fn(high, shape, *a, **k)
                calls.append((int(high), int(shape), res.clone().numpy().astype(np.int64)))
                return res
            return wrapped

        ref2d.grid_monte_carlo_sample = recording(originals[0])
        ref2d.grid_as_monte_carlo_sample = recording(originals[1])
        try:
            new_keys, loss = ref2d.compute_contra_memobank_loss(
                rep_all, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], memobank, ptrs, caps,
                rep_teacher.detach(), delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries,
                num_negatives=spec.negatives, temp=spec.temp)[-2:]
        finally:
            ref2d.grid_monte_carlo_sample, ref2d.grid_as_monte_carlo_sample = originals
        loss.backward()
        p = f"s{step}_"
        out[p + "loss"] = loss.detach().numpy()
        out[p + "new_keys"] = np.asarray(new_keys, np.int64)
        out[p + "call_high"] = np.asarray([c[0] for c in calls], np.int64)
        out[p + "call_shape"] = np.asarray([c[1] for c in calls], np.int64)
        for k, c in enumerate(calls):
            out[p + f"call{k}"] = c[2].astype(np.int32)
        for c in range(spec.classes):
            out[p + f"bank{c}"] = memobank[c][0].detach().numpy().copy()
        out[p + "ptrs"] = np.asarray([int(t[0]) for t in ptrs], np.int64)
        # the producers' outputs (strided samples) and every gradient the step leaves behind
        out[p + "rep_sample"] = rep_all.detach()[:, ::8, ::4, ::4].numpy().copy()
        out[p + "rep_teacher_sample"] = rep_teacher.detach()[:, ::8, ::4, ::4].numpy().copy()
        out[p + "grad_map4_l"] = maps_l[4].grad.numpy().copy()
        out[p + "grad_map4_u"] = maps_u[4].grad.numpy().copy()
        out[p + "grad_map0_u"] = maps_u[0].grad.numpy().copy()
        ones = torch.ones(rep_all.shape[1])
        for name, w in (("fea4", q_fe.fea4.weight), ("qrep0", q_rep[0].weight), ("qrep1", q_rep[1].weight), ("fea3", q_fe.fea3.weight)):
            g = w.grad.reshape(w.shape[0], w.shape[1])
            out[p + f"gw_{name}_rowsum"] = g.sum(dim=1).numpy().copy()
            out[p + f"gw_{name}_colsum"] = g.sum(dim=0).numpy().copy()
            out[p + f"gw_{name}_norm"] = np.float32(g.norm())
    return out


def run_3d():
    """model_3D.FeatureExtractor_3d (model_3D.py:20-63) with the 3-D trainer's channel plan and q_representation
    (train_arco_3d.py:206-213): output samples for the twin module."""
    lines = open("/root/reference/code/model_3D.py").read().split("\n")
    d0 = next(i for i, ln in enumerate(lines) if ln.startswith("class FeatureExtractor_3d("))
    d1 = next(i for i in range(d0 + 1, len(lines)) if lines[i] and not lines[i][0].isspace())
    ns = dict(torch=torch, nn=nn, F=F, np=np)
    exec(compile("\n".join(lines[d0:d1]), "<model_3D.FeatureExtractor_3d>", "exec"), ns)
    tl = open("/root/reference/code/train_arco_3d.py").read().split("\n")
    a = next(i for i, ln in enumerate(tl) if ln.strip().startswith("q_representation = nn.Sequential("))
    b = next(i for i in range(a, len(tl)) if tl[i].strip().startswith("q_feature_extractor = FeatureExtractor_3d("))
    exec(compile(textwrap.dedent("\n".join(tl[a:b + 1])), "<train_arco_3d.py:206-213>", "exec"), ns)
    q_rep, q_fe = ns["q_representation"], ns["q_feature_extractor"]
    x = producer_inputs_3d()
    with torch.no_grad():
        for i in range(5):
            getattr(q_fe, f"fea{i}").weight.copy_(x["w_fe"][i].view_as(getattr(q_fe, f"fea{i}").weight))
        q_rep[0].weight.copy_(x["w_rep"][0].view_as(q_rep[0].weight))
        q_rep[1].weight.copy_(x["w_rep"][1].view_as(q_rep[1].weight))
        fea = q_fe(x["maps"])
        rep = q_rep(fea)
    return dict(fea=fea.numpy().copy(), rep=rep.numpy().copy())


def main():
    import warnings
    warnings.filterwarnings("ignore")
    for case in PRODUCER_CASES:
        res = run(case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **res)
        spec = case["spec"]
        print(case["name"], [float(res[f"s{t}_loss"]) for t in range(spec.steps)],
              [res[f"s{t}_new_keys"].tolist() for t in range(spec.steps)], f"{os.path.getsize(path) / 1024:.0f} KB")


    res = run_3d()
    path = os.path.join(HERE, "producers_3d.npz")
    np.savez_compressed(path, **res)
    print("producers_3d", res["rep"].shape, f"{os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
