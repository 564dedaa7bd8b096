"""Parity cases shared by the golden-vector generator and the tests.

Every case is tiny (the real reference finishes it in well under a second on one core) and
targets one of the edge cases SURVEY.md section 8(c) asks for.  Inputs come from
``arco_b200.synth.exact_case`` so they are bit-identical on every machine.
"""
from arco_b200.synth import CaseSpec

CASES = [
    # plain 2-D, fresh zero-row bank as train_arco_2d.py:147-154 builds it; two calls so the bank grows
    CaseSpec("acdc_smc", 2, 2, 4, (24, 24), 8, queries=16, negatives=8, func="smc", steps=2),
    # antithetic sampler, non-square image, D not a multiple of 8, randn row as train_arco_3d.py:148
    CaseSpec("acdc_asmc", 2, 2, 4, (20, 28), 12, queries=32, negatives=6, func="asmc",
             bank_init="randn1", steps=2, seed=2024),
    # grid-sampler path on both calls (anchor candidates and bank length >= 57)
    CaseSpec("grid_paths", 3, 3, 4, (32, 32), 8, queries=64, negatives=4, func="smc",
             bank_init="fill:100", caps=[150, 120, 120, 120], seed=99),
    CaseSpec("grid_paths_as", 3, 3, 4, (32, 32), 8, queries=64, negatives=4, func="asmc",
             bank_init="fill:100", caps=[150, 120, 120, 120], seed=98),
    # 19 classes: exercises the rank window [3,20) with C > 16
    CaseSpec("city19", 2, 2, 19, (16, 16), 16, queries=8, negatives=16, func="smc",
             bank_init="fill:40", caps=[64] * 19, steps=2, seed=7),
    # 3-D volume twin (loss_helper.py); C=2 so no key is ever enqueued (trap 3)
    CaseSpec("la3d", 1, 1, 2, (8, 8, 6), 16, queries=16, negatives=8, func="asmc",
             bank_init="randn1", steps=2, seed=11),
    # class 1 absent: list-position / class-id mismatch (trap 1)
    CaseSpec("absent_class", 2, 2, 4, (16, 16), 8, func="smc", bank_init="fill:30",
             caps=[50, 30, 30, 30], label_mode="absent:1", seed=5),
    # only one class present: zero loss that still depends on rep (trap 5)
    CaseSpec("single_class", 2, 2, 4, (12, 12), 8, func="smc", label_mode="single:2", seed=6),
    # class 1 has no confident pixel: empty anchor list, slot skipped but counted (trap 2)
    CaseSpec("no_anchor", 2, 2, 4, (16, 16), 8, func="smc", bank_init="fill:20",
             caps=[40, 40, 40, 40], label_mode="noanchor:1", seed=8),
    # more keys in one call than the queue holds: eviction keeps the newest rows (trap 6)
    CaseSpec("overflow", 2, 2, 5, (16, 16), 8, func="smc", bank_init="fill:4",
             caps=[7, 5, 5, 5, 5], mask_frac=0.9, steps=3, seed=13),
    # any other func string: plain torch.randint (loss_helper_3d.py:335-338)
    CaseSpec("uniform_func", 2, 2, 4, (16, 16), 8, func="rand", bank_init="fill:25",
             caps=[40, 40, 40, 40], seed=17),
    # half of the unlabelled pixels carry the ignore label -1 -> class 0 one-hot (trap 4)
    CaseSpec("ignore_heavy", 2, 2, 4, (16, 16), 8, func="smc", bank_init="fill:25",
             caps=[40, 40, 40, 40], ignore_frac=0.5, seed=19),
    # blocky labels, unequal labelled / unlabelled split, image size not a multiple of 4
    CaseSpec("blocky_odd", 1, 3, 6, (15, 13), 20, queries=24, negatives=5, func="asmc",
             bank_init="fill:60", caps=[80] * 6, label_mode="blocky", seed=23),
    # optional EMA prototypes (a11, loss_helper_3d.py:488-497): non-zero momentum with class 1 absent, and the all-zero
    # momentum tensor that leaves the prototypes untouched; 3-tuple return
    CaseSpec("momentum", 2, 2, 4, (16, 16), 8, func="smc", bank_init="fill:30", caps=[50, 30, 30, 30],
             label_mode="absent:1", momentum="rand", i_iter=7, seed=37),
    CaseSpec("momentum_zero", 2, 2, 4, (16, 16), 8, func="asmc", bank_init="fill:30", caps=[50, 30, 30, 30],
             momentum="zeros", i_iter=3, seed=41),
    # far more queries than anchor candidates: every candidate is drawn many times, backward must accumulate (trap 8)
    CaseSpec("dup_anchors", 1, 1, 4, (8, 8), 8, queries=64, negatives=4, func="smc", bank_init="fill:12",
             caps=[20, 20, 20, 20], seed=43),
    # bf16 representation tensors (config 2); compared at bf16 tolerance
    CaseSpec("bf16_rep", 2, 2, 4, (16, 16), 16, func="smc", bank_init="fill:25",
             caps=[40, 40, 40, 40], dtype="bf16", seed=29),
]

BY_NAME = {c.name: c for c in CASES}

# (func, high, shape, seed) -> reference sampler output, stored in tests/golden/samplers.npz
SAMPLER_CASES = [
    (func, high, shape, seed)
    for func in ("smc", "asmc")
    for (high, shape, seed) in [
        (1, 256, 1), (5, 16, 2), (15, 256, 3), (16, 256, 4), (17, 31, 5), (40, 7, 6), (56, 256, 7),
        (57, 256, 8), (64, 256, 9), (100, 256, 10), (1000, 256, 11), (6133, 256, 12),
        (300, 2048, 13), (29929, 4096, 14), (30000, 8192, 15), (30000, 131072, 16),
        (3000, 1, 17), (5000, 20, 18),
    ]
]


# SURVEY.md section 8(f) rank 1 -- mask / threshold preparation (train_arco_2d.py:345-393, train_arco_3d.py:315-353):
# (name, n_lab, n_unlab, classes, spatial, alpha_t, ignore_frac, logit quantisation step (0 = none), seed)
PREPARE_CASES = [
    ("prep2d_a14", 2, 2, 4, (24, 20), 14.0, 0.05, 0.0, 101),
    ("prep2d_a20", 1, 3, 4, (16, 16), 20.0, 0.0, 0.0, 102),
    ("prep2d_ties", 2, 2, 4, (32, 32), 7.4, 0.10, 0.5, 103),       # coarse logits -> many exactly equal entropies
    ("prep2d_c19", 1, 2, 19, (12, 12), 3.0, 0.05, 0.0, 104),
    ("prep3d_c2", 1, 2, 2, (8, 8, 6), 11.2, 0.05, 0.0, 105),
    ("prep2d_a0", 1, 1, 4, (8, 8), 0.0, 0.0, 0.0, 106),           # last epoch: alpha_t = 0 -> thresholds = min / max
]


def prepare_inputs(case):
    """Machine-independent inputs of a PREPARE case (integers scaled to floats, no libm involved)."""
    import numpy as np
    import torch
    name, n_lab, n_unlab, C, spatial, alpha_t, ign, quant, seed = case
    rs = np.random.RandomState(seed)

    def logits(b):
        x = rs.randint(-4096, 4096, size=(b, C) + tuple(spatial)).astype(np.float32) / 1024.0
        if quant:
            x = np.round(x / quant) * quant
        return torch.from_numpy(x.astype(np.float32))

    def labels(b):
        lab = rs.randint(0, C, size=(b,) + tuple(spatial)).astype(np.int64)
        if ign:
            lab[rs.rand(*lab.shape) < ign] = -1
        return torch.from_numpy(lab)

    return dict(pred_l=logits(n_lab), pred_u=logits(n_unlab), pred_l_teacher=logits(n_lab), pred_u_teacher=logits(n_unlab),
                train_l_label=labels(n_lab), train_u_aug_label=labels(n_unlab), alpha_t=float(alpha_t), num_classes=C)


# SURVEY.md section 8(f) rank 3 -- revisiting loss + random-pool queue (train_arco_2d.py:126-136, :109-120, :400-402)
REVISIT_CASES = [
    dict(name="revisit_small", bs=4, K=8, shape=(8, 8, 8), topk=3, steps=3, seed=201, dtype="f32"),
    dict(name="revisit_ref_tiles", bs=12, K=36, shape=(6, 8, 10), topk=5, steps=2, seed=202, dtype="f32"),     # the trainers' bs / K / topk
    dict(name="revisit_ragged", bs=3, K=9, shape=(5, 4, 12), topk=5, steps=4, seed=203, dtype="f32"),          # L = 240: one partial chunk
    dict(name="revisit_bf16", bs=6, K=12, shape=(16, 8, 8), topk=4, steps=2, seed=204, dtype="bf16"),
]


def revisit_inputs(case):
    """Inputs of a REVISIT case: per-step student / teacher tensors [bs, D, H, W] and the initial unit-row pool [K, L]."""
    import numpy as np
    import torch
    rs = np.random.RandomState(case["seed"])
    bs, K = case["bs"], case["K"]
    D, H, W = case["shape"]
    L = D * H * W
    pool = rs.randint(-512, 513, size=(K, L)).astype(np.float64)
    pool[:, 0] += 0.5                                             # never an all-zero row
    pool = (pool / np.sqrt((pool * pool).sum(axis=1, keepdims=True))).astype(np.float32)
    tdt = torch.bfloat16 if case["dtype"] == "bf16" else torch.float32

    def rep():
        base = rs.randint(-256, 257, size=(bs, D, H, W)).astype(np.float32) / np.float32(64.0)
        return torch.from_numpy(base).to(tdt).float()            # bf16 cases: bf16-representable values, handed over as fp32

    reps_s, reps_t = [], []
    for _ in range(case["steps"]):
        s = rep()
        # teacher = student + a perturbation, plus a pull towards one pool row so that neighbours are not arbitrary
        t = (s * 0.75 + rep() * 0.25)
        pull = torch.from_numpy(pool[rs.randint(0, K, size=bs)]).view(bs, D, H, W) * float(np.sqrt(L)) * 0.5
        reps_s.append((s + pull).to(tdt).float())
        reps_t.append((t + pull).to(tdt).float())
    return dict(pool=torch.from_numpy(pool), rep_u=reps_s, rep_u_teacher=reps_t)


# SURVEY.md section 8(f) rank 4 -- compute_unsupervised_loss (train_arco_2d.py:482-489)
UNSUP_CASES = [
    dict(name="unsup_small", B=2, C=4, spatial=(12, 12), strong_threshold=0.97, ignore_frac=0.1, seed=301),
    dict(name="unsup_c19", B=3, C=19, spatial=(9, 14), strong_threshold=0.9, ignore_frac=0.3, seed=302),
    dict(name="unsup_noignore", B=1, C=2, spatial=(16, 16), strong_threshold=0.5, ignore_frac=0.0, seed=303),
]


def unsup_inputs(case):
    import numpy as np
    import torch
    rs = np.random.RandomState(case["seed"])
    B, C, sp = case["B"], case["C"], tuple(case["spatial"])
    predict = torch.from_numpy(rs.randint(-4096, 4096, size=(B, C) + sp).astype(np.float32) / np.float32(1024.0))
    target = rs.randint(0, C, size=(B,) + sp).astype(np.int64)
    if case["ignore_frac"]:
        target[rs.rand(*target.shape) < case["ignore_frac"]] = -1
    logits = torch.from_numpy(rs.randint(0, 1025, size=(B,) + sp).astype(np.float32) / np.float32(1024.0))
    return dict(predict=predict, target=torch.from_numpy(target), logits=logits)


# SURVEY.md section 8(f) rank 4 -- TPS equivariance loss (train_arco_2d.py:404-423, tps/rand_tps.py:82-153)
EQV_CASES = [
    dict(name="eqv_small", B=4, C=4, H=32, W=32, sigma=0.05, weak_threshold=0.7, seed=401),
    dict(name="eqv_rect", B=2, C=3, H=24, W=40, sigma=0.1, weak_threshold=0.5, seed=402),
    dict(name="eqv_c19", B=2, C=19, H=16, W=16, sigma=0.05, weak_threshold=0.7, seed=403),
]


def eqv_inputs(case):
    import numpy as np
    import torch
    rs = np.random.RandomState(case["seed"] + 7)
    B, C, H, W = case["B"], case["C"], case["H"], case["W"]

    def f(*shape, lo=-4096, hi=4096, div=1024.0):
        return torch.from_numpy(rs.randint(lo, hi, size=shape).astype(np.float32) / np.float32(div))

    return dict(images=f(B, 1, H, W, lo=0, hi=1025), pred_all=f(B, C, H, W), pred_tps=f(B, C, H, W),
                labels=torch.from_numpy(rs.randint(0, C, size=(B, H, W)).astype(np.int64)), logits=f(B, H, W, lo=0, hi=1025))


# SURVEY.md section 8(f) rank 2 -- representation producers (model_2D.py:20-55, train_arco_2d.py:231-236, :313-333)
# The trainers' own channel plan [256,128,64,32,16] -> D = 496 on a small pyramid (1x1 ... 16x16 = 256 pixels per image).
PRODUCER_CASES = [
    # two steps: the second one wraps the class rings (fill 40 of 60/50 rows, ~32 keys per class and step)
    dict(name="producers_d496", fea_dim=(256, 128, 64, 32, 16), top=16, spec=CaseSpec(
        "producers_d496", 2, 2, 4, (16, 16), 496, queries=16, negatives=8, func="smc", bank_init="fill:40",
        caps=[60, 50, 50, 50], steps=2, seed=601)),
]


def producer_inputs(case, step=0):
    """Machine-independent inputs of a PRODUCER case: the five decoder feature maps of the labelled / unlabelled batch for the
    student and the teacher network (coarsest first, like the reference's ``fea_list``), and the weights of the two
    FeatureExtractors (``fea0..fea4``) and of ``q_representation``; labels / probabilities / masks come from
    ``exact_case(case['spec'], step)``.  Integer RNG + exact IEEE scaling only."""
    import numpy as np
    import torch
    spec = case["spec"]
    rs = np.random.RandomState(spec.seed * 31 + 17)               # weights: the same at every step
    fea = list(case["fea_dim"])

    def tri(shape, scale):
        a = rs.randint(-256, 257, size=shape).astype(np.int32) + rs.randint(-256, 257, size=shape).astype(np.int32)
        return torch.from_numpy((a.astype(np.float32) * np.float32(scale)).astype(np.float32))

    def extractor_weights():
        ws, cnt = [], 0
        for i in range(5):
            cnt += fea[i]
            # ~ N(0, 1/cin): keeps the activations O(1) through the residual chain
            ws.append(tri((cnt, cnt), 2.0 ** -9 / 2 ** int(np.log2(cnt) / 2 - 3)))
        return ws

    w_q_fe, w_k_fe = extractor_weights(), extractor_weights()
    D = sum(fea)
    w_q_rep = [tri((D, D), 2.0 ** -10), tri((D, D), 2.0 ** -10)]
    rs2 = np.random.RandomState(spec.seed * 31 + 18 + 101 * step)
    top = case["top"]
    sizes = [max(1, top >> (4 - i)) for i in range(5)]

    def maps(batch):
        return [torch.from_numpy((rs2.randint(-256, 257, size=(batch, fea[i], sizes[i], sizes[i])).astype(np.float32)
                                  / np.float32(128.0)).astype(np.float32)) for i in range(5)]

    return dict(w_q_fe=w_q_fe, w_k_fe=w_k_fe, w_q_rep=w_q_rep,
                maps_l=maps(spec.n_lab), maps_u=maps(spec.n_unlab),
                maps_l_teacher=maps(spec.n_lab), maps_u_teacher=maps(spec.n_unlab))


def producer_inputs_3d():
    """Inputs of the 3-D producers twin check: the 3-D trainer's channel plan [128, 64, 32, 16, 16] -> 256 -> output_dim 16
    (train_arco_3d.py:212-213) on a tiny pyramid (1 .. 8 voxels per axis), exact integer-derived values."""
    import numpy as np
    import torch
    rs = np.random.RandomState(733)
    fea = [128, 64, 32, 16, 16]

    def tri(shape, scale):
        a = rs.randint(-256, 257, size=shape).astype(np.int32) + rs.randint(-256, 257, size=shape).astype(np.int32)
        return torch.from_numpy((a.astype(np.float32) * np.float32(scale)).astype(np.float32))

    ws, cnt = [], 0
    for i in range(5):
        cnt += fea[i]
        out = cnt if i < 4 else 16
        ws.append(tri((out, cnt), 2.0 ** -12))
    w_rep = [tri((16, 16), 2.0 ** -9), tri((16, 16), 2.0 ** -9)]
    sizes = [1, 2, 4, 8, 8]
    maps = [torch.from_numpy((rs.randint(-256, 257, size=(2, fea[i], sizes[i], sizes[i], sizes[i])).astype(np.float32)
                              / np.float32(128.0)).astype(np.float32)) for i in range(5)]
    return dict(w_fe=ws, w_rep=w_rep, maps=maps)
