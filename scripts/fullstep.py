"""BASELINE.json config 2: one synthetic ARCO 2-D training step on 1 B200 -- U-Net student + EMA teacher,
FeatureExtractor([256,128,64,32,16] -> 496) + q_representation (2 x Conv1x1), supervised CE + Dice, and every
semi-supervised term of the reference's step -- run twice on identical weights and data:

  (a) "reference_ops": the loss terms as the reference composes them from ATen ops (the oracle restatements of
      compute_contra_memobank_loss with its CPU memory bank and samplers, the mask preparation with its two
      np.percentile round trips, get_revisiting_loss + pool enqueue, compute_unsupervised_loss, RandTPS + the
      equivariance block), executed with CUDA tensors;
  (b) "arco_b200": the same step with those terms replaced by this repository's CUDA ops.

The backbones are PyTorch in both arms (north_star: "the U-Net/V-Net/DeepLab backbones stay in PyTorch"): a plain
re-statement of the reference's 2-D U-Net (networks/unetWithArgs.py:29-160: two 3x3 conv + BN + LeakyReLU per level,
16..256 channels, transposed-conv decoder that also returns its five feature maps) and of FeatureExtractor's 1x1-conv
cascade (model_2D.py:20-55).  Follows train_arco_2d.py:284-431; augmentation (PIL / CutMix, CPU side) is replaced by
synthetic tensors, bf16 autocast.  Imported by bench.py (--workload acdc2d_fullstep); also runnable on its own.
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CH = [16, 32, 64, 128, 256]
DROP = [0.05, 0.1, 0.2, 0.3, 0.5]
REP = sum(CH)                                                     # 496


def block(cin, cout, p):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU(), nn.Dropout(p),
                         nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU())


class UNet2D(nn.Module):
    """networks/unetWithArgs.py:309-348 (Encoder :77-107, Decoder :109-160): returns (logits, bottleneck, [5 decoder maps])."""

    def __init__(self, in_ch=1, classes=4):
        super().__init__()
        self.inc = block(in_ch, CH[0], DROP[0])
        self.down = nn.ModuleList([nn.Sequential(nn.MaxPool2d(2), block(CH[i], CH[i + 1], DROP[i + 1])) for i in range(4)])
        self.up = nn.ModuleList([nn.ConvTranspose2d(CH[4 - i], CH[3 - i], 2, stride=2) for i in range(4)])
        self.upc = nn.ModuleList([block(2 * CH[3 - i], CH[3 - i], 0.0) for i in range(4)])
        self.out = nn.Conv2d(CH[0], classes, 3, padding=1)

    def forward(self, x):
        enc = [self.inc(x)]
        for d in self.down:
            enc.append(d(enc[-1]))
        y = enc[4]
        maps = [y]
        for i in range(4):
            y = self.upc[i](torch.cat([enc[3 - i], self.up[i](y)], dim=1))
            maps.append(y)
        return self.out(y), enc[4], maps


class FeatureFuse(nn.Module):
    """model_2D.py:20-55: per level x = conv1x1(x) + x, bilinear upsample (align_corners) to the next map, concat; last
    level conv1x1 -> 496 channels at full resolution."""

    def __init__(self):
        super().__init__()
        dims = [CH[4], CH[3], CH[2], CH[1], CH[0]]
        cnt, convs = 0, []
        for i, d in enumerate(dims):
            cnt += d
            convs.append(nn.Conv2d(cnt, cnt if i < 4 else REP, 1, bias=False))
        self.convs = nn.ModuleList(convs)

    def forward(self, maps):
        x = self.convs[0](maps[0]) + maps[0]
        for i in range(1, 5):
            x = F.interpolate(x, size=maps[i].shape[-2:], mode="bilinear", align_corners=True)
            x = torch.cat((x, maps[i]), dim=1)
            x = self.convs[i](x) + x if i < 4 else self.convs[i](x)
        return x


def dice_loss(prob, target, classes):
    """utils/losses.py:173-209 (DiceLoss, softmax already applied)."""
    loss = 0.0
    for i in range(classes):
        t = (target == i).float()
        s = prob[:, i]
        inter, y, z = torch.sum(s * t), torch.sum(t * t), torch.sum(s * s)
        loss = loss + (1 - (2 * inter + 1e-5) / (z + y + 1e-5))
    return loss / classes


class Step:
    def __init__(self, impl, dev, n_lab=12, n_unlab=12, classes=4, size=256, seed=1337, pool_rows=36):
        assert impl in ("reference_ops", "arco_b200")
        self.impl, self.dev, self.C, self.size, self.nl, self.nu = impl, dev, classes, size, n_lab, n_unlab
        torch.manual_seed(seed)
        random.seed(seed)
        np.random.seed(seed)
        self.model, self.ema = UNet2D(1, classes).to(dev), UNet2D(1, classes).to(dev)
        self.ema.load_state_dict(self.model.state_dict())
        for p in self.ema.parameters():
            p.requires_grad_(False)
        self.q_rep = nn.Sequential(nn.Conv2d(REP, REP, 1, bias=False), nn.Conv2d(REP, REP, 1, bias=False)).to(dev)   # :231-234
        self.q_fe, self.k_fe = FeatureFuse().to(dev), FeatureFuse().to(dev)                                          # :236-237
        self.k_fe.load_state_dict(self.q_fe.state_dict())
        for p in self.k_fe.parameters():
            p.requires_grad_(False)                                                                                  # :250-253
        params = [p for p in self.model.parameters()] + list(self.q_rep.parameters()) + list(self.q_fe.parameters())
        self.opt = torch.optim.SGD(params, lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True)                  # :248
        g = torch.Generator(device=dev).manual_seed(seed)
        self.x_l = torch.rand(n_lab, 1, size, size, device=dev, generator=g)
        self.x_u = torch.rand(n_unlab, 1, size, size, device=dev, generator=g)
        self.x_l2 = torch.rand(n_lab, 1, size, size, device=dev, generator=g)          # images_cj2_l
        self.x_u2 = torch.rand(n_unlab, 1, size, size, device=dev, generator=g)        # images_cj2_u
        self.y_l = torch.randint(0, classes, (n_lab, size, size), device=dev, generator=g)
        # memory bank / pools exactly as train_arco_2d.py:147-159 builds them
        self.memobank = [[torch.zeros(1, REP)] for _ in range(classes)]
        self.queue_size = [50000] + [30000] * (classes - 1)
        self.queue_ptr = [torch.zeros(1, dtype=torch.long) for _ in range(classes)]
        pool = torch.randn(pool_rows, REP * size * size, device=dev, generator=g)
        self.pool = F.normalize(pool, dim=1)
        self.pool_ptr = torch.zeros(1, dtype=torch.long)
        if impl == "arco_b200":
            import arco_b200
            self.ops = arco_b200
            self.tps = arco_b200.RandTPS(size, size, batch_size=n_lab + n_unlab, sigma=0.01, random_scale=(0.8, 1.2), mode="affine")
        else:
            import oracle
            from oracle import prepare_oracle
            self.ops, self.prep = oracle, prepare_oracle
            from arco_b200.stepterms import draw_source_control_points     # host RNG restatement (CPU); grid by the oracle
            self._draw = draw_source_control_points
            self._ctrl = torch.Tensor([(a, b) for a in torch.arange(-1.0, 1.00001, 0.5) for b in torch.arange(-1.0, 1.00001, 0.5)])
        self.iter = 0
        self.t_terms = 0.0

    # --- the semi-supervised terms, two implementations ------------------------------------------------------------
    def _terms_arco(self, d):
        o = self.ops
        prep = o.prepare_contrast_inputs(d["pred_u"].float(), d["pred_l_t"].float(), d["pred_u_t"].float(), self.y_l, d["y_u"], d["alpha_t"])
        reco = o.compute_contra_memobank_loss(
            d["rep_all"], prep["label_l"], prep["label_u"], prep["prob_l_teacher"], prep["prob_u_teacher"], prep["low_mask_all"],
            prep["high_mask_all"], self.memobank, self.queue_ptr, self.queue_size, d["rep_all_t"], delta_n=0.97, func="smc",
            num_queries=256, num_negatives=512, sparse_grad=True)[-1]
        loss_q = o.get_revisiting_loss(self.pool, d["rep_u"], d["rep_u_t"], topk=5)
        o.revisit_enqueue(d["rep_u_t"], self.pool, self.pool_ptr)
        unsup = o.compute_unsupervised_loss(d["pred_u"].float(), d["y_u"], d["conf_u"], 0.97)
        self.tps.reset_control_points()
        images_tps = self.tps(torch.cat((self.x_l2, self.x_u2)))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred_tps = self.model(images_tps)[0]
        eqv = o.tps_equivariance_loss(pred_tps.float(), d["pred_all"].detach().float(), self.tps, torch.cat((self.y_l, d["y_u"])),
                                      torch.cat((torch.full_like(d["conf_u"][: self.nl], 255.0), d["conf_u"])), 0.7)
        return reco, loss_q, unsup, eqv

    def _terms_reference(self, d):
        o = self.ops
        prep = self.prep.prepare(d["pred_u"].float(), d["pred_l_t"].float(), d["pred_u_t"].float(), self.y_l, d["y_u"], d["alpha_t"], self.C)
        sampler = o.grid_strata_sample
        reco = o.contra_memobank_loss(
            d["rep_all"].float(), prep["label_l"].to(self.dev).long(), prep["label_u"].to(self.dev).long(), prep["prob_l_teacher"],
            prep["prob_u_teacher"], prep["low_mask_all"].to(self.dev), prep["high_mask_all"].to(self.dev), self.memobank, self.queue_ptr,
            self.queue_size, d["rep_all_t"].float(), delta_n=0.97, sampler=sampler, num_queries=256, num_negatives=512).loss
        loss_q = o.revisiting_loss(self.pool, d["rep_u"], d["rep_u_t"], topk=5)[0]
        o.pool_enqueue(d["rep_u_t"], self.pool, self.pool_ptr)
        unsup = o.unsupervised_loss(d["pred_u"].float(), d["y_u"], d["conf_u"], 0.97)
        B = self.nl + self.nu
        src = self._draw(self._ctrl, B, 0.01, (1.0 / 1.2, 1.0 / 0.8), "affine", True)
        grid = o.tps_grid(src, self.size, self.size).to(self.dev)
        images_tps = o.warp(torch.cat((self.x_l2, self.x_u2)), grid)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred_tps = self.model(images_tps)[0]
        eqv = o.equivariance_loss(pred_tps.float(), d["pred_all"].detach().float(), grid, torch.cat((self.y_l, d["y_u"])),
                                  torch.cat((torch.full_like(d["conf_u"][: self.nl], 255.0), d["conf_u"])), 0.7)[0]
        return reco, loss_q, unsup, eqv

    def step(self):
        dev = self.dev
        self.iter += 1
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            p = self.ema(self.x_u)[0]                                                           # :284-286
        conf_u, y_u = torch.max(torch.softmax(p.float(), dim=1), dim=1)
        with torch.no_grad():                                                                   # EMA of the key extractor (:306-308)
            for pq, pk in zip(self.q_fe.parameters(), self.k_fe.parameters()):
                pk.data.mul_(0.99).add_(pq.data, alpha=0.01)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred_l, _, fm_l = self.model(self.x_l)                                              # :310-312
            _, _, fm_l2 = self.model(self.x_l2)
            pred_u, _, fm_u = self.model(self.x_u)
            with torch.no_grad():
                pred_l_t, _, fm_l_t = self.ema(self.x_l)                                        # :314-315
                pred_u_t, _, fm_u_t = self.ema(self.x_u)
                rep_l_t, rep_u_t = self.k_fe(fm_l_t), self.k_fe(fm_u_t)                         # :321-322
            rep_u = self.q_rep(self.q_fe(fm_u))                                                 # :317-326
            rep_l = self.q_rep(self.q_fe(fm_l))
            _ = self.q_rep(self.q_fe(fm_l2))
            rep_all, pred_all = torch.cat((rep_l, rep_u)), torch.cat((pred_l, pred_u))
            rep_all_t = torch.cat((rep_l_t, rep_u_t))
        prob_l = torch.softmax(pred_l.float(), dim=1)
        sup = F.cross_entropy(pred_l.float(), self.y_l) + dice_loss(prob_l, self.y_l, self.C)   # :336-339
        d = dict(pred_u=pred_u, pred_l_t=pred_l_t, pred_u_t=pred_u_t, y_u=y_u, conf_u=conf_u, alpha_t=20.0 * (1 - 0.3),
                 rep_all=rep_all, rep_all_t=rep_all_t.detach(), rep_u=rep_u.detach(), rep_u_t=rep_u_t.detach(), pred_all=pred_all)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        reco, loss_q, unsup, eqv = (self._terms_arco if self.impl == "arco_b200" else self._terms_reference)(d)
        torch.cuda.synchronize(dev)
        self.t_terms += time.perf_counter() - t0
        loss = 0.01 * reco + 1.0 * unsup + sup + 1.0 * eqv + 1.0 * loss_q                       # :426 (k1 = 0.01)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        with torch.no_grad():                                                                   # isd._momentum_update_key_encoder (:431)
            for ps, pt in zip(self.model.parameters(), self.ema.parameters()):
                pt.data.mul_(0.99).add_(ps.data, alpha=0.01)
        return loss.detach(), (reco.detach(), loss_q.detach(), unsup.detach(), eqv.detach())


def run(impl, steps=5, warmup=2, dev=None, **kw):
    dev = dev or torch.device("cuda", 0)
    st = Step(impl, dev, **kw)
    for _ in range(warmup):
        st.step()
    torch.cuda.synchronize(dev)
    st.t_terms = 0.0
    t0 = time.perf_counter()
    last = None
    for _ in range(steps):
        last = st.step()
    torch.cuda.synchronize(dev)
    ms = (time.perf_counter() - t0) / steps * 1e3
    P = (st.nl + st.nu) * st.size * st.size
    out = dict(impl=impl, ms_per_step=ms, ms_semi_supervised_terms=st.t_terms / steps * 1e3, steps=steps, warmup=warmup,
               value_mpixels_per_s=P / (ms * 1e-3) / 1e6, loss=float(last[0]),
               terms=dict(zip(("reco", "loss_q", "unsup", "eqv"), (float(v) for v in last[1]))),
               peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
    del st
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    for impl in (sys.argv[2:] or ["arco_b200", "reference_ops"]):
        print(json.dumps(run(impl, steps=steps)), flush=True)
