"""SURVEY.md section 8(f) rank 2 -- the representation producers, folded into the contrastive loss.

Reference (``/root/reference/code``):

* ``model_2D.py:20-55``  ``FeatureExtractor``: four residual bias-free 1x1 convolutions interleaved with bilinear
  up-sampling and channel concatenation, then ``fea4`` -- a plain bias-free 1x1 convolution ``[496 -> 496]`` at full
  resolution that PRODUCES the representation tensor;
* ``train_arco_2d.py:231-234``  ``q_representation``: two more bias-free 1x1 convolutions on the student side;
* ``train_arco_2d.py:317-333``  ``rep = q_representation(q_fe(student maps))``, ``rep_teacher = k_fe(teacher maps)``, both
  ``[24, 496, 256, 256]``, handed to ``compute_contra_memobank_loss`` (``:394-398``).

The loss consumes of those two tensors only per-class SUMS of ``rep_teacher``, its rows at the key pixels, and the rows of
``rep`` at the <= C*Q sampled anchors.  A bias-free 1x1 convolution is linear and per-pixel, so it commutes with all three
selections (``sum_px W x = W sum_px x``, ``(W x)[p] = W x[p]``).  :func:`compute_contra_memobank_loss_from_features` therefore
takes the INPUTS of ``fea4`` (what :meth:`FeatureExtractor.trunk` returns) plus the weights, runs the one-pass
classify / prototype / enqueue kernels on those inputs and applies the weights afterwards, only where a value is read:
the enqueued ring rows by an in-place tcgen05 GEMM (``arco_keys_transform``), the class sums in fp64
(``arco_proto_transform``), the anchors as ``[C*Q, D]`` rows (``arco_anchor_gather`` -> three small matrix products ->
``arco_infonce_rows``).  Neither ``rep`` nor ``rep_teacher`` is materialised; the values (loss, gradients w.r.t. the student
features AND the three student weights, ring rows, ``new_keys``) are those of the reference composition up to fp32
summation order.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from .bank import DeviceMemoryBank
from .contra import (DELTA_P, HIGH_RANK, LOW_RANK, LazyKeys, _FUNC, _GEOMETRY, _PREFILL_GRAD, _PREFILL_MIN_BYTES, _sampler_stream,
                     _side_stream, _sparse_state)


class FeatureExtractor(nn.Module):
    """Twin of ``model_2D.FeatureExtractor`` (``model_2D.py:20-55``): same constructor, same parameter names
    (``fea0.weight`` .. ``fea4.weight``, so a reference ``state_dict`` loads unchanged), same ``forward``.
    :meth:`trunk` is ``forward`` without the last convolution -- the tensor
    :func:`compute_contra_memobank_loss_from_features` takes."""

    def __init__(self, fea_dim=(256, 128, 64, 32, 16), output_dim=256) -> None:
        super().__init__()
        fea_dim = list(fea_dim)
        if len(fea_dim) != 5:
            raise AssertionError("input_dim is not correct")
        cnt = fea_dim[0]
        self.fea0 = nn.Conv2d(cnt, cnt, kernel_size=1, bias=False)
        cnt += fea_dim[1]
        self.fea1 = nn.Conv2d(cnt, cnt, kernel_size=1, bias=False)
        cnt += fea_dim[2]
        self.fea2 = nn.Conv2d(cnt, cnt, kernel_size=1, bias=False)
        cnt += fea_dim[3]
        self.fea3 = nn.Conv2d(cnt, cnt, kernel_size=1, bias=False)
        cnt += fea_dim[4]
        self.fea4 = nn.Conv2d(cnt, output_dim, kernel_size=1, bias=False)

    def trunk(self, fea_list):
        """``model_2D.py:36-52``: everything up to (not including) ``fea4``."""
        f0, f1, f2, f3, f4 = fea_list[:5]
        up = lambda t, ref: F.interpolate(t, size=ref.shape[-2:], mode="bilinear", align_corners=True)
        x = self.fea0(f0) + f0
        x = torch.cat((up(x, f1), f1), dim=1)
        x = self.fea1(x) + x
        x = torch.cat((up(x, f2), f2), dim=1)
        x = self.fea2(x) + x
        x = torch.cat((up(x, f3), f3), dim=1)
        x = self.fea3(x) + x
        return torch.cat((up(x, f4), f4), dim=1)

    def forward(self, fea_list):
        return self.fea4(self.trunk(fea_list))


class FeatureExtractor_3d(nn.Module):
    """Twin of ``model_3D.FeatureExtractor_3d`` (``model_3D.py:20-63``, used by ``train_arco_3d.py:212-213``): the same chain with
    ``Conv3d`` / trilinear up-sampling.  Its last convolution is ``[output_dim = 16] x [256]`` -- NOT square -- so folding it into
    the loss would make the prototype pass read the 256-channel input instead of the 16-channel output; that does not pay
    (and with C = 2 no key is ever enqueued, trap 3), so the 3-D producers stay a module: ``forward`` only."""

    def __init__(self, fea_dim=(128, 64, 32, 16, 16), output_dim=128) -> None:
        super().__init__()
        fea_dim = list(fea_dim)
        if len(fea_dim) != 5:
            raise AssertionError("input_dim is not correct")
        cnt = fea_dim[0]
        self.fea0 = nn.Conv3d(cnt, cnt, kernel_size=1, bias=False)
        cnt += fea_dim[1]
        self.fea1 = nn.Conv3d(cnt, cnt, kernel_size=1, bias=False)
        cnt += fea_dim[2]
        self.fea2 = nn.Conv3d(cnt, cnt, kernel_size=1, bias=False)
        cnt += fea_dim[3]
        self.fea3 = nn.Conv3d(cnt, cnt, kernel_size=1, bias=False)
        cnt += fea_dim[4]
        self.fea4 = nn.Conv3d(cnt, output_dim, kernel_size=1, bias=False)

    def trunk(self, fea_list):
        f0, f1, f2, f3, f4 = fea_list[:5]
        up = lambda t, ref: F.interpolate(t, size=ref.shape[-3:], mode="trilinear", align_corners=True)
        x = self.fea0(f0) + f0
        x = torch.cat((up(x, f1), f1), dim=1)
        x = self.fea1(x) + x
        x = torch.cat((up(x, f2), f2), dim=1)
        x = self.fea2(x) + x
        x = torch.cat((up(x, f3), f3), dim=1)
        x = self.fea3(x) + x
        return torch.cat((up(x, f4), f4), dim=1)

    def forward(self, fea_list):
        return self.fea4(self.trunk(fea_list))


def make_q_representation(dim: int = 256 + 128 + 64 + 32 + 16) -> nn.Sequential:
    """``train_arco_2d.py:231-234``."""
    return nn.Sequential(nn.Conv2d(dim, dim, kernel_size=1, bias=False), nn.Conv2d(dim, dim, kernel_size=1, bias=False))


def _w2d(w: torch.Tensor, D: int, name: str) -> torch.Tensor:
    if w.dim() in (4, 5) and all(int(k) == 1 for k in w.shape[2:]):       # Conv2d / Conv3d 1x1(x1) weight
        w = w.reshape(w.shape[0], w.shape[1])
    if tuple(w.shape) != (D, D):
        raise ValueError(f"{name} must be a bias-free 1x1 convolution weight [{D},{D}] or [{D},{D},1,1], got {tuple(w.shape)}")
    return w


class _GatherRows(torch.autograd.Function):
    """rows = x[:, :, anchor pixels] (already gathered by ``arco_anchor_gather``); backward scatters the row gradients
    into a dense (or, opt-in, op-owned sparse) gradient of ``x``."""

    @staticmethod
    def forward(ctx, x, rows, pix, st):
        ctx.st = st
        ctx.pix = pix
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        return rows.view_as(rows)

    @staticmethod
    def backward(ctx, grad_rows):
        st = ctx.st
        dev = grad_rows.device
        sp = torch.cuda.current_stream(dev).cuda_stream
        g = grad_rows.detach().to(torch.float32).contiguous()
        one = torch.ones(1, dtype=torch.float32, device=dev)
        d = C.byref(st["dims"])
        if st["sparse"] is not None:
            grad_x, prev_pix = st["sparse"]
            _cabi.check(_cabi.lib.arco_grad_scatter_sparse(d, g.data_ptr(), ctx.pix.data_ptr(), one.data_ptr(),
                                                           grad_x.data_ptr(), prev_pix.data_ptr(), sp), "arco_grad_scatter_sparse")
        elif st.get("grad_buf") is not None:
            # zero-filled during forward on the side stream (underneath the read-bound forward kernels): scatter only
            grad_x = st["grad_buf"]
            st["grad_buf"] = None                            # a second backward (retain_graph) takes the slow path
            _cabi.check(_cabi.lib.arco_grad_scatter_add(d, g.data_ptr(), ctx.pix.data_ptr(), one.data_ptr(), grad_x.data_ptr(), sp),
                        "arco_grad_scatter_add")
        else:
            grad_x = torch.empty(ctx.x_shape, dtype=ctx.x_dtype, device=dev)
            _cabi.check(_cabi.lib.arco_grad_scatter(d, g.data_ptr(), ctx.pix.data_ptr(), one.data_ptr(), grad_x.data_ptr(), sp),
                        "arco_grad_scatter")
        return grad_x, None, None, None


class _InfoNCERows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a_rows, st):
        dims, bank = st["dims"], st["bank"]
        dev = a_rows.device
        sp = torch.cuda.current_stream(dev).cuda_stream
        Cn, Q, N, D = dims.classes, dims.queries, dims.negatives, dims.feat
        a = a_rows.detach().to(torch.float32).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g_anchor = torch.empty((Cn * Q, D), dtype=torch.float32, device=dev)
        pix_out = torch.empty((Cn, Q), dtype=torch.int32, device=dev)
        debug = st["debug"]
        logits = torch.zeros((Cn, Q, 1 + N), dtype=torch.float32, device=dev) if debug is not None else None
        _cabi.check(_cabi.lib.arco_infonce_rows(
            C.byref(dims), a.data_ptr(), st["pix"].data_ptr(), C.byref(bank.c_struct), st["proto_sums"].data_ptr(),
            st["idx_a"].data_ptr(), st["idx_n"].data_ptr(), float(st["temp"]), loss.data_ptr(), g_anchor.data_ptr(),
            pix_out.data_ptr(), logits.data_ptr() if logits is not None else None, st["ws"].data_ptr(), sp), "arco_infonce_rows")
        if debug is not None:
            debug.update(logits=logits, grad_anchor=g_anchor.view(Cn, Q, D), anchor_rows=a.view(Cn, Q, D))
        ctx.save_for_backward(g_anchor)
        ctx.in_dtype = a_rows.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (g_anchor,) = ctx.saved_tensors
        return (g_anchor * grad_out.to(torch.float32)).to(ctx.in_dtype), None


def compute_contra_memobank_loss_from_features(
    x_student: torch.Tensor,
    x_teacher: torch.Tensor,
    student_weights: Sequence[torch.Tensor],
    teacher_weight: torch.Tensor,
    label_l,
    label_u,
    prob_l,
    prob_u,
    low_mask,
    high_mask,
    memobank,
    queue_prtlis,
    queue_size,
    delta_n=1.0,
    func="asmc",
    num_queries=256,
    num_negatives=512,
    temp=0.5,
    *,
    process_group=None,
    seed: Optional[int] = None,
    sparse_grad: bool = False,
    _inject: Optional[dict] = None,
    _debug: Optional[dict] = None,
):
    """``compute_contra_memobank_loss(rep, ..., rep_teacher, ...)`` (``loss_helper_3d.py:271-513``) for

        rep         = conv1x1(... conv1x1(x_student, student_weights[0]) ..., student_weights[-1])
        rep_teacher = conv1x1(x_teacher, teacher_weight)

    i.e. ``train_arco_2d.py:317-333`` with ``x_* = FeatureExtractor.trunk(maps)``, ``student_weights = [q_fe.fea4.weight,
    q_representation[0].weight, q_representation[1].weight]`` and ``teacher_weight = k_fe.fea4.weight`` -- without ever
    forming ``rep`` / ``rep_teacher``.  Weights are ``[D, D]`` (or ``[D, D, 1, 1]``); ``x_*`` are ``[B, D, *spatial]`` fp32 or
    bf16 (bf16: every product is rounded to bf16 like an autocast convolution's output).  Labels / probabilities / masks /
    memory bank and the returned ``(new_keys, loss)`` are those of :func:`arco_b200.compute_contra_memobank_loss`;
    ``loss.backward()`` delivers the gradient of ``x_student`` (dense; ``sparse_grad=True``: the op-owned buffer contract of
    the plain op) and of every student weight that requires grad.  The ``momentum_prototype`` branch is not offered here.
    """
    if not (torch.is_tensor(x_student) and x_student.is_cuda):
        raise RuntimeError("arco_b200.producers needs CUDA tensors: there is no CPU fallback")
    if x_student.dim() not in (4, 5):
        raise ValueError(f"x_student must be [B,D,H,W] or [B,D,H,W,Z], got {tuple(x_student.shape)}")
    if x_student.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError(f"features must be float32 or bfloat16, got {x_student.dtype}")
    if x_teacher.shape != x_student.shape or x_teacher.dtype != x_student.dtype or x_teacher.device != x_student.device:
        raise ValueError("x_teacher must match x_student in shape, dtype and device")
    dev = x_student.device
    B, D = int(x_student.shape[0]), int(x_student.shape[1])
    spatial = tuple(x_student.shape[2:])
    S = 1
    for s in spatial:
        S *= int(s)
    n_lab, n_unlab = int(label_l.shape[0]), int(label_u.shape[0])
    if n_lab + n_unlab != B:
        raise ValueError(f"label_l ({n_lab}) + label_u ({n_unlab}) images != batch ({B})")
    Cn = int(prob_l.shape[1]) if prob_l.numel() else int(prob_u.shape[1])
    if label_l.dim() == x_student.dim():
        label_kind = _cabi.LABEL_ONEHOT_I64
    elif label_l.dim() == x_student.dim() - 1:
        label_kind = _cabi.LABEL_INDEX_I64
    else:
        raise ValueError("label_l must be one-hot [B_l,C,*S] or an integer map [B_l,*S]")
    if label_l.dtype != torch.int64 or label_u.dtype != torch.int64:
        raise ValueError("labels must be int64")
    for name, t, n in (("prob_l", prob_l, n_lab), ("prob_u", prob_u, n_unlab)):
        if t.dtype != torch.float32 or tuple(t.shape) != (n, Cn) + spatial:
            raise ValueError(f"{name} must be float32 [{n},{Cn},*spatial], got {t.dtype} {tuple(t.shape)}")
    for name, t in (("low_mask", low_mask), ("high_mask", high_mask)):
        if t.dtype != torch.float32 or tuple(t.shape) != (B, 1) + spatial:
            raise ValueError(f"{name} must be float32 [B,1,*spatial], got {t.dtype} {tuple(t.shape)}")
    if Cn > _cabi.MAX_CLASSES or len(memobank) != Cn:
        raise ValueError(f"memobank has {len(memobank)} classes, prob has {Cn} (max {_cabi.MAX_CLASSES})")
    if D % 8 != 0 or D < 16 or D > 512:
        raise ValueError(f"feature size D must be a multiple of 8 in [16, 512], got {D}")
    if B * S >= 2 ** 31:
        raise ValueError("more than 2^31 pixels per call")
    if len(student_weights) < 1:
        raise ValueError("student_weights must hold at least one 1x1 convolution weight")
    w_s = [_w2d(w, D, f"student_weights[{i}]") for i, w in enumerate(student_weights)]
    w_k = _w2d(teacher_weight.detach(), D, "teacher_weight")
    cdt = x_student.dtype

    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        sp = stream.cuda_stream
        lib = _cabi.lib
        bank = DeviceMemoryBank.adopt(memobank, queue_prtlis, queue_size, D, dev, cdt)
        bank.poll()
        bank.begin_step()
        sampler_seed, sampler_step = _sampler_stream(seed, bank, dev)
        key = (n_lab, n_unlab, Cn, D, S, int(num_queries), int(num_negatives), cdt, label_kind, dev.index)
        cached = _GEOMETRY.get(key)
        if cached is None:
            dims = _cabi.Dims(n_lab, n_unlab, Cn, D, S, int(num_queries), int(num_negatives),
                              _cabi.BF16 if cdt == torch.bfloat16 else _cabi.F32, label_kind)
            cached = _GEOMETRY[key] = (dims, _cabi.workspace_layout(dims))
        dims, layout = cached
        Q, N = dims.queries, dims.negatives
        d, b = C.byref(dims), C.byref(bank.c_struct)
        ws = torch.empty(layout.total_bytes, dtype=torch.uint8, device=dev)
        wsp = ws.data_ptr()
        xs = x_student.detach().contiguous()
        xt = x_teacher.detach().contiguous()
        hold = [t.contiguous() for t in (label_l, label_u, prob_l, prob_u, low_mask, high_mask)]
        grad_buf, fill_side = None, None
        if (x_student.requires_grad and torch.is_grad_enabled() and not sparse_grad and _PREFILL_GRAD
                and x_student.numel() * x_student.element_size() >= _PREFILL_MIN_BYTES):
            # the dense gradient of x_student must be zero-filled whatever the inputs are: start that now on a side stream
            fill_side = _side_stream(dev, 1)
            grad_buf = torch.empty(x_student.shape, dtype=cdt, device=dev)
            fill_side.wait_stream(stream)
            _cabi.check(lib.arco_grad_zero(d, grad_buf.data_ptr(), fill_side.cuda_stream), "arco_grad_zero")
        ptr = lambda t, n: t.data_ptr() if n else None
        _cabi.check(lib.arco_classify_plan(d, ptr(hold[0], n_lab), ptr(hold[1], n_unlab), ptr(hold[2], n_lab), ptr(hold[3], n_unlab),
                                           hold[4].data_ptr(), hold[5].data_ptr(), DELTA_P, float(delta_n), LOW_RANK, HIGH_RANK,
                                           b, wsp, sp), "arco_classify_plan")
        idx_a = torch.empty((Cn, Q), dtype=torch.int32, device=dev)
        idx_n = torch.empty((Cn, Q * max(N, 1)), dtype=torch.int32, device=dev)
        func_id = _FUNC.get(func, _cabi.FUNC_UNIFORM)
        side = None
        if _inject is None:
            side = _side_stream(dev)                         # the sampler only needs the plan: under the prototype pass
            side.wait_stream(stream)
            _cabi.check(lib.arco_sample(d, func_id, sampler_seed, sampler_step, idx_a.data_ptr(), idx_n.data_ptr(), wsp,
                                        side.cuda_stream), "arco_sample")
        # prototype sums and key rows of the convolution's INPUT (same kernels, same bytes as for rep_teacher itself)
        proto_x = torch.empty((Cn, D + 1), dtype=torch.float64, device=dev)
        _cabi.check(lib.arco_proto_enqueue(d, xt.data_ptr(), b, proto_x.data_ptr(), wsp, sp), "arco_proto_enqueue")
        # ... then the teacher's weight, only where a value is consumed
        ring_f32 = bank.c_struct.row_dtype == _cabi.F32
        wk = w_k.to(cdt)                                     # the precision the reference convolution multiplies in
        wk_ring = (wk.float() if ring_f32 else wk.to(torch.bfloat16)).contiguous()
        scratch = torch.empty(max(1, lib.arco_keys_transform_scratch_bytes(D, bank.c_struct.row_dtype)), dtype=torch.uint8, device=dev)
        _cabi.check(lib.arco_keys_transform(d, b, wk_ring.data_ptr(), scratch.data_ptr(), wsp, sp), "arco_keys_transform")
        if process_group is not None:
            torch.distributed.all_reduce(proto_x, group=process_group)       # the one exchange step, in x space
        proto_sums = torch.empty_like(proto_x)
        _cabi.check(lib.arco_proto_transform(Cn, D, wk_ring.data_ptr(), _cabi.F32 if ring_f32 else _cabi.BF16,
                                             proto_x.data_ptr(), proto_sums.data_ptr(), sp), "arco_proto_transform")
        if side is not None:
            stream.wait_stream(side)
        if process_group is not None:
            _cabi.check(lib.arco_replan_global(d, proto_sums.data_ptr(), wsp, sp), "arco_replan_global")
            if side is not None:
                _cabi.check(lib.arco_sample_if_replanned(d, func_id, sampler_seed, sampler_step, idx_a.data_ptr(),
                                                         idx_n.data_ptr(), wsp, sp), "arco_sample_if_replanned")
        plan_view = ws[layout.plan: layout.plan + C.sizeof(_cabi.Plan)]
        if _inject is not None:
            plan = _cabi.Plan.from_buffer_copy(plan_view.cpu().numpy().tobytes())
            active = [j for j in range(Cn) if plan.slot_active[j]]
            if len(_inject["anchor"]) != len(active) or len(_inject["neg"]) != len(active):
                raise ValueError(f"_inject carries {len(_inject['anchor'])} index sets, the step has {len(active)} active positions")
            idx_a.zero_()
            idx_n.zero_()
            for k, j in enumerate(active):
                idx_a[j] = _inject["anchor"][k].to(dev, torch.int32)
                idx_n[j, : Q * N] = _inject["neg"][k].to(dev, torch.int32)
        # anchors as rows of the student's convolution input
        rows_x = torch.empty((Cn * Q, D), dtype=torch.float32, device=dev)
        pix = torch.empty((Cn * Q,), dtype=torch.int32, device=dev)
        _cabi.check(lib.arco_anchor_gather(d, xs.data_ptr(), idx_a.data_ptr(), rows_x.data_ptr(), pix.data_ptr(), wsp, sp),
                    "arco_anchor_gather")
        needs_grad = torch.is_grad_enabled() and (x_student.requires_grad or any(w.requires_grad for w in w_s))
        state = dict(dims=dims, bank=bank, ws=ws, proto_sums=proto_sums, idx_a=idx_a, idx_n=idx_n, pix=pix, temp=temp,
                     debug=_debug, grad_buf=grad_buf,
                     sparse=_sparse_state(bank, x_student, Cn * Q) if (sparse_grad and x_student.requires_grad
                                                                        and torch.is_grad_enabled()) else None)
        a = _GatherRows.apply(x_student, rows_x, pix, state) if (needs_grad and x_student.requires_grad) else rows_x
        a = a.to(cdt)
        for w in w_s:                                        # [C*Q, D] x [D, D]: plain library products, autograd's
            a = a @ w.to(cdt).t()
        loss = _InfoNCERows.apply(a, state)
        if fill_side is not None:
            stream.wait_stream(fill_side)                   # from here on the buffer is ordinary main-stream memory
        bank.post_step(plan_view)
        if _debug is not None:
            _debug.update(ws=ws, layout=layout, dims=dims, proto_sums=proto_sums, proto_sums_x=proto_x, anchor_pix=pix.view(Cn, Q),
                          idx_anchor=idx_a, idx_neg=idx_n, rows_x=rows_x.view(Cn, Q, D))
    return LazyKeys(bank, Cn, plan_view), loss
