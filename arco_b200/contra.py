"""Drop-in ``compute_contra_memobank_loss`` backed by hand-written sm_100a kernels.

Mirrors the reference operator (same name, positional order, keyword names, return tuple and side
effects): ``/root/reference/code/loss_helper_3d.py:271-513`` for ``rep [B,D,H,W]`` and
``/root/reference/code/loss_helper.py:442-686`` for ``rep [B,D,H,W,Z]``; call sites
``train_arco_2d.py:394-398`` / ``train_arco_3d.py:356-360``.  One implementation serves both: the
spatial axes are flattened to ``S`` and never permuted.

Pipeline (all on ``torch.cuda.current_stream()``, no host synchronisation in forward or backward):

  arco_classify_plan (classify + scans + plan) -> arco_proto_enqueue (+ fp64 finalize) -> [exchange of the
  C x (D+1) fp64 prototype sums] -> arco_sample (or injected indices) -> arco_infonce        (forward)
  arco_grad_scatter                                                                 (backward)

There is no CPU path and no PyTorch fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from typing import List, Optional

import torch

from . import _cabi
from .bank import DeviceMemoryBank

DELTA_P = 0.3                  # current_class_threshold, loss_helper_3d.py:316
LOW_RANK, HIGH_RANK = 3, 20    # loss_helper_3d.py:318

_FUNC = {"smc": _cabi.FUNC_SMC, "asmc": _cabi.FUNC_ASMC}
_GEOMETRY = {}                 # problem shape -> (arco_dims, workspace layout)
_PREFILL_GRAD = os.environ.get("ARCO_PREFILL_GRAD", "1") != "0"   # zero-fill grad_rep during forward (side stream)
# measured: pays from ~250 MB of rep (la3d 0.232 -> 0.218 ms), costs at 134 MB (acdc2d_loss 0.155 -> 0.181 ms: the extra side-stream
# launches outweigh a 25 us fill)
_PREFILL_MIN_BYTES = int(os.environ.get("ARCO_PREFILL_MIN_MB", "192")) << 20


class LazyKeys(list):
    """``new_keys`` (reference: list of ints, loss_helper_3d.py:404-411) resolved on first access -- a device->host read of
    the step's own ``arco_plan`` -- so the step itself never waits for the device."""

    def __init__(self, bank: DeviceMemoryBank, classes: int, plan_view: torch.Tensor):
        super().__init__()
        self._bank = bank
        self._classes = classes
        self._plan_view = plan_view
        self._done = False

    def _resolve(self):
        if not self._done:
            plan = _cabi.Plan.from_buffer_copy(self._plan_view.cpu().numpy().tobytes())
            self._plan_view = None
            self._bank.check_status(plan.status)
            super().extend(int(plan.n_key[c]) for c in range(self._classes))
            self._done = True

    def __getitem__(self, i):
        self._resolve()
        return super().__getitem__(i)

    def __iter__(self):
        self._resolve()
        return super().__iter__()

    def __len__(self):
        return self._classes

    def __eq__(self, other):
        self._resolve()
        return list(self) == list(other)

    def __repr__(self):
        self._resolve()
        return super().__repr__()


_SIDE_STREAMS = {}


def _side_stream(dev: torch.device, which: int = 0) -> torch.cuda.Stream:
    """Per-device helper streams: 0 = sampler (joins before InfoNCE), 1 = grad_rep zero fill (joins in backward)."""
    st = _SIDE_STREAMS.get((dev.index, which))
    if st is None:
        st = _SIDE_STREAMS[(dev.index, which)] = torch.cuda.Stream(device=dev)
    return st


_P2P = {}                # (group id, device index, n) -> peer-mapped exchange buffer state, or None when unavailable
P2P_EXCHANGE_USED = False
EXCHANGE_PLANE = None    # "nvlink-p2p" | "nccl": which data plane the last multi-GPU step used (also logged once per group)
_log = logging.getLogger("arco_b200")


_XCHG_STEP_WORD = 1 << 63      # arco_exchange.seq bit 63: the exchange buffer carries a step word behind its flags


def _agree(flag: bool, group, dev: torch.device) -> bool:
    """True iff ``flag`` is true on EVERY rank (one MIN all-reduce)."""
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def _p2p_exchange(group, dev: torch.device, n: int):
    """Symmetric (peer-mapped) buffer for the one exchange step of the multi-GPU path, created once per
    (process group, device, C*(D+1)).  Returns None -- the caller then uses an NCCL all-reduce -- when torch's symmetric
    memory cannot map the peers (no NVLink/P2P, older torch) or ``ARCO_P2P_ALLREDUCE=0``.

    The decision is a collective and is taken in two agreed phases so that no rank can be left alone inside a
    rendezvous: (1) every rank reports whether it CAN try (env switch, symmetric-memory module importable, every peer
    device reachable with ``can_device_access_peer``) and only if all can do they enter ``symm.empty``/``rendezvous``;
    (2) every rank reports whether the mapping succeeded.  Only the expected failure types are caught."""
    global EXCHANGE_PLANE
    key = (getattr(group, "group_name", None) or id(group), dev.index, n)
    if key in _P2P:
        return _P2P[key]
    world = torch.distributed.get_world_size(group)
    rank = torch.distributed.get_rank(group)
    symm, why = None, ""
    if os.environ.get("ARCO_P2P_ALLREDUCE", "1") == "0":
        why = "ARCO_P2P_ALLREDUCE=0"
    else:
        try:
            import torch.distributed._symmetric_memory as symm          # noqa: WPS433
        except ImportError as e:
            why = f"torch symmetric memory unavailable ({e})"
        if symm is not None:
            ndev = torch.cuda.device_count()
            if not all(torch.cuda.can_device_access_peer(dev.index, o) for o in range(ndev) if o != dev.index):
                symm, why = None, "a peer GPU is not P2P-accessible"
    state = None
    if _agree(symm is not None, group, dev):
        try:
            slot = (n + 63) // 64 * 64
            buf = symm.empty(2 * slot + 64 + 8, dtype=torch.float64, device=dev)     # two slots, 64 flags, the step word
            buf.zero_()
            hdl = symm.rendezvous(buf, group)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) == world and all(ptrs):
                torch.cuda.synchronize(dev)
                torch.distributed.barrier(group)            # every rank's flags are zero before anyone signals
                state = dict(buf=buf, hdl=hdl, rank=rank, world=world, slot=slot, seq=0, seq_flags=_XCHG_STEP_WORD,
                             peers=torch.tensor(ptrs, dtype=torch.int64, device=dev))
            else:
                why = "rendezvous returned an incomplete peer pointer table"
        except (RuntimeError, ValueError, AttributeError, TypeError) as e:
            why = f"symmetric-memory rendezvous failed: {e}"
        if not _agree(state is not None, group, dev):
            state = None
            why = why or "a peer could not map the exchange buffer"
    elif not why:
        why = "another rank cannot use peer memory"
    EXCHANGE_PLANE = "nvlink-p2p" if state is not None else "nccl"
    if rank == 0:
        _log.info("arco_b200 multi-GPU exchange (C*(D+1)=%d fp64, world %d): %s%s", n, world,
                  "own kernel over NVLink peer memory" if state is not None else "NCCL all-reduce",
                  "" if state is not None else f" ({why})")
    _P2P[key] = state
    return state


_BANK_SERIAL = 0          # banks that took the default-seed path, in creation order (deterministic across restarts)


def _sampler_stream(seed, bank: DeviceMemoryBank, dev: torch.device):
    """(Philox seed, per-step stream id) of the in-kernel sampler.

    The kernels add the bank's DEVICE step counter to the stream id returned here, so even a CUDA-graph replay of
    identical launch parameters draws a fresh stream every step.
    Explicit ``seed=``: deterministic replay -- stream id = number of steps this bank has completed (tests, benchmarks).
    Default: behave like the reference, which CONSUMES global RNG state (Python ``random`` + torch's CPU generator,
    loss_helper_3d.py:157-177): the seed is torch's CUDA seed mixed with the process's distributed rank (DDP ranks seeded
    alike must not draw identical index streams), and the stream id starts at -- and every step advances -- the device's
    default CUDA generator offset, so a resumed run that restores (or re-seeds) torch's RNG state continues (or
    replays) exactly like the reference would, instead of restarting at stream 0 whenever a bank is re-adopted."""
    if seed is not None:
        return int(seed) & (2 ** 64 - 1), 0              # the device adds the bank's own step counter (arco_plan.step_ctr)
    rank = torch.distributed.get_rank() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
    gen = torch.cuda.default_generators[dev.index]
    off = int(gen.get_offset())
    gen.set_offset(off + 4)                                  # one Philox counter block per step, never reused
    s0 = (int(gen.initial_seed()) ^ ((rank + 1) * 0x9E3779B97F4A7C15)) & (2 ** 63 - 1)
    # The stream id handed to the kernels is the generator offset AT THE BANK'S FIRST STEP; the device adds the bank's step
    # counter.  Every later step still advances the generator (so the offset a checkpoint saves keeps growing and a resumed
    # run's first offset lies beyond every stream this run used), but the launch parameter stays constant from step to
    # step -- which is what lets arco_forward replay its launch sequence as a CUDA graph (forward.cu).
    # Banks that live side by side (two loss heads) get different Philox keys: their stream ids overlap by construction.
    base = bank.__dict__.get("_stream_base")
    if base is None or base[0] != s0:
        global _BANK_SERIAL
        _BANK_SERIAL += 1
        base = bank.__dict__["_stream_base"] = (s0, off // 4, (s0 ^ (_BANK_SERIAL * 0xD1B54A32D192ED03)) & (2 ** 63 - 1))
    return base[2], base[1]


def _sparse_state(bank: DeviceMemoryBank, rep: torch.Tensor, rows: int):
    """(grad_rep buffer, previous anchor pixels) of the opt-in sparse-gradient contract, kept per bank and shape."""
    key = (tuple(rep.shape), rep.dtype, rows)
    cache = bank.__dict__.setdefault("_sparse_grad", {})
    st = cache.get(key)
    if st is None:
        cache.clear()                                        # one live shape per bank: do not pin several dense buffers
        st = cache[key] = (torch.zeros(rep.shape, dtype=rep.dtype, device=rep.device),
                           torch.full((rows,), -1, dtype=torch.int32, device=rep.device))
    return st


def _flat(t: torch.Tensor, lead: int) -> torch.Tensor:
    """Contiguous view with the trailing spatial axes flattened (``lead`` leading axes kept)."""
    t = t.contiguous()
    return t.view(*t.shape[:lead], -1)


class _ContraLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rep, st):
        dims, bank = st["dims"], st["bank"]
        dev = rep.device
        stream = torch.cuda.current_stream(dev)
        sp = stream.cuda_stream
        lib = _cabi.lib
        layout = st["layout"]
        if st.get("logits") is None and st["inject"] is None and (st["debug"] is None or st["debug"].get("fused")):
            # one FFI call; multi-GPU too when the exchange buffer could be peer-mapped (no NCCL call between the stages)
            p2p = _p2p_exchange(st["group"], dev, dims.classes * (dims.feat + 1)) if st["group"] is not None else None
            if st["group"] is None or p2p is not None:
                return _ContraLoss._forward_fused(ctx, rep, st, dims, bank, layout, dev, sp, p2p)
        ws = torch.empty(layout.total_bytes, dtype=torch.uint8, device=dev)
        wsp = ws.data_ptr()
        Cn, Q, N, D = dims.classes, dims.queries, dims.negatives, dims.feat
        d = C.byref(dims)
        b = C.byref(bank.c_struct)
        if st["prefill"]:
            # The dense grad_rep must be zero-filled whatever the inputs are (autograd contract, P*D*e bytes of HBM
            # writes).  Start that fill now on the side stream: it runs underneath the forward kernels, which are
            # read-bound, and backward only has to scatter.  The buffer comes from the main stream's pool.
            side0 = _side_stream(dev, 1)
            buf = torch.empty(rep.shape, dtype=rep.dtype, device=dev)
            side0.wait_stream(stream)                       # the block may still be in use by earlier main-stream work
            _cabi.check(lib.arco_grad_zero(d, buf.data_ptr(), side0.cuda_stream), "arco_grad_zero")
            st["grad_buf"], st["side"] = buf, side0

        lg = st.get("logits")
        if lg is not None:
            # logits-in: student entropy -> two percentile thresholds (device), then classify reads the teacher LOGITS, the label
            # maps and the entropy directly; no probability / mask tensor is ever written (arco_classify_plan_logits)
            n_u_px = dims.n_unlab * dims.space
            ent = torch.empty(max(n_u_px, 1), dtype=torch.float32, device=dev)
            thr = torch.full((2,), float("nan"), dtype=torch.float32, device=dev)
            if dims.n_unlab:
                _cabi.check(lib.arco_softmax_rows(lg["pred_u"].data_ptr(), dims.n_unlab, Cn, dims.space, None, ent.data_ptr(), sp),
                            "arco_softmax_rows")
                scratch = torch.empty(int(lib.arco_entropy_masks_scratch()), dtype=torch.uint8, device=dev)
                _cabi.check(lib.arco_entropy_thresholds(ent.data_ptr(), st["label_u"].data_ptr(), n_u_px, lg["q_low"], lg["q_high"],
                                                        thr.data_ptr(), scratch.data_ptr(), sp), "arco_entropy_thresholds")
            _cabi.check(lib.arco_classify_plan_logits(
                d, st["label_l"].data_ptr() if st["label_l"] is not None else None,
                st["label_u"].data_ptr() if st["label_u"] is not None else None,
                lg["pred_l_teacher"].data_ptr() if dims.n_lab else None, lg["pred_u_teacher"].data_ptr() if dims.n_unlab else None,
                ent.data_ptr(), thr.data_ptr(), DELTA_P, float(st["delta_n"]), LOW_RANK, HIGH_RANK, b, wsp, sp),
                "arco_classify_plan_logits")
            st["thresholds"] = thr
        ins = (st["label_l"].data_ptr() if st["label_l"] is not None else None,
               st["label_u"].data_ptr() if st["label_u"] is not None else None,
               st["prob_l"].data_ptr() if st["prob_l"] is not None else None,
               st["prob_u"].data_ptr() if st["prob_u"] is not None else None,
               st["low_mask"].data_ptr() if st["low_mask"] is not None else None,
               st["high_mask"].data_ptr() if st["high_mask"] is not None else None,
               DELTA_P, float(st["delta_n"]), LOW_RANK, HIGH_RANK)
        if lg is not None:
            pass
        elif st["debug"] is not None and st["debug"].get("legacy_scan"):
            # the two-launch form of the C ABI (memset + classify, then the stand-alone scan/plan kernel)
            _cabi.check(lib.arco_classify_count(d, *ins, wsp, sp), "arco_classify_count")
            _cabi.check(lib.arco_scan_plan(d, b, wsp, sp), "arco_scan_plan")
        else:
            _cabi.check(lib.arco_classify_plan(d, *ins, b, wsp, sp), "arco_classify_plan")
        proto_sums = torch.empty((Cn, D + 1), dtype=torch.float64, device=dev)
        p2p = _p2p_exchange(st["group"], dev, Cn * (D + 1)) if st["group"] is not None else None
        proto_local = proto_sums                            # where the prototype kernel writes this rank's sums
        if p2p is not None:
            p2p["seq"] += 1
            off = (p2p["seq"] & 1) * p2p["slot"]
            proto_local = p2p["buf"][off: off + Cn * (D + 1)]
        idx_a = torch.empty((Cn, Q), dtype=torch.int32, device=dev)
        idx_n = torch.empty((Cn, Q * max(N, 1)), dtype=torch.int32, device=dev)
        plan_view = ws[layout.plan: layout.plan + C.sizeof(_cabi.Plan)]
        group, inject = st["group"], st["inject"]
        side = None
        if inject is None:
            # the sampler only needs the plan: run it on a side stream underneath the prototype pass (multi-GPU: on the
            # rank-local plan, speculatively -- redone below only if the global valid-class list changes the plan)
            side = _side_stream(dev)
            side.wait_stream(stream)
            _cabi.check(lib.arco_sample(d, st["func"], st["seed"], st["step"], idx_a.data_ptr(), idx_n.data_ptr(),
                                        wsp, side.cuda_stream), "arco_sample")
        _cabi.check(lib.arco_proto_enqueue(d, st["rep_teacher"].data_ptr(), b, proto_local.data_ptr(), wsp, sp),
                    "arco_proto_enqueue")
        if group is not None:
            # the one exchange step of the path: C*(D+1) fp64 sums + counts
            if p2p is not None:
                global P2P_EXCHANGE_USED
                P2P_EXCHANGE_USED = True
                _cabi.check(lib.arco_proto_allreduce_p2p(d, p2p["peers"].data_ptr(), p2p["rank"], p2p["world"],
                                                         p2p["seq"] | p2p.get("seq_flags", 0),
                                                         p2p["slot"], proto_sums.data_ptr(), wsp, sp),
                            "arco_proto_allreduce_p2p")
            else:
                torch.distributed.all_reduce(proto_sums, group=group)
            if side is not None:
                stream.wait_stream(side)                    # the speculative sampler has read the local plan
            _cabi.check(lib.arco_replan_global(d, proto_sums.data_ptr(), wsp, sp), "arco_replan_global")
            if side is not None:
                _cabi.check(lib.arco_sample_if_replanned(d, st["func"], st["seed"], st["step"], idx_a.data_ptr(),
                                                         idx_n.data_ptr(), wsp, sp), "arco_sample_if_replanned")
        elif side is not None:
            stream.wait_stream(side)
        if inject is not None:
            # parity tests: replay the reference's own indices, one (anchor, negative) pair per active position
            plan = _cabi.Plan.from_buffer_copy(plan_view.cpu().numpy().tobytes())
            active = [j for j in range(Cn) if plan.slot_active[j]]
            if len(inject["anchor"]) != len(active) or len(inject["neg"]) != len(active):
                raise ValueError(f"_inject carries {len(inject['anchor'])} index sets, the step has {len(active)} "
                                 f"active positions {active}")
            idx_a.zero_()
            idx_n.zero_()
            for k, j in enumerate(active):
                idx_a[j] = inject["anchor"][k].to(dev, torch.int32)
                idx_n[j, : Q * N] = inject["neg"][k].to(dev, torch.int32)

        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g_anchor = torch.empty((Cn, Q, D), dtype=torch.float32, device=dev)
        pix = torch.empty((Cn, Q), dtype=torch.int32, device=dev)
        debug = st["debug"]
        logits = torch.zeros((Cn, Q, 1 + N), dtype=torch.float32, device=dev) if debug is not None else None
        mom = st["momentum"]
        if mom is None:
            _cabi.check(lib.arco_infonce(
                d, st["rep_data"].data_ptr(), b, proto_sums.data_ptr(), idx_a.data_ptr(), idx_n.data_ptr(),
                float(st["temp"]), loss.data_ptr(), g_anchor.data_ptr(), pix.data_ptr(),
                logits.data_ptr() if logits is not None else None, wsp, sp), "arco_infonce")
        else:
            # a11: EMA prototypes.  The "is the momentum tensor all zero" test (:489) and the <=1-valid-class early
            # return (:421-424, which hands back the input tensor) are both resolved on the device.
            mom_on = (mom != 0).any().to(torch.int32).reshape(1)
            proto_out = torch.zeros_like(mom)
            _cabi.check(lib.arco_infonce_ema(
                d, st["rep_data"].data_ptr(), b, proto_sums.data_ptr(), idx_a.data_ptr(), idx_n.data_ptr(),
                float(st["temp"]), loss.data_ptr(), g_anchor.data_ptr(), pix.data_ptr(),
                logits.data_ptr() if logits is not None else None, mom.data_ptr(), mom_on.data_ptr(),
                float(st["ema_decay"]), float(1.0 - st["ema_decay"]), proto_out.data_ptr(), wsp, sp), "arco_infonce_ema")
            n_valid = ws[layout.plan + 384: layout.plan + 388].view(torch.int32)      # arco_plan.n_valid
            st["prototype_out"] = torch.where(n_valid <= 1, mom, proto_out)
        bank.post_step(plan_view)
        st["plan_view"] = plan_view
        ctx.fused = None
        if debug is not None:
            debug.update(ws=ws, layout=layout, dims=dims, proto_sums=proto_sums, logits=logits, anchor_pix=pix,
                         grad_anchor=g_anchor, idx_anchor=idx_a, idx_neg=idx_n)
        ctx.save_for_backward(g_anchor, pix)
        ctx.dims = dims
        ctx.rep_shape = rep.shape
        ctx.rep_dtype = rep.dtype
        ctx.prefilled = None
        ctx.sparse = st["sparse"]
        if st["prefill"]:
            # The dense grad_rep must be zero-filled whatever the inputs are (autograd contract, P*D*e bytes of
            # HBM writes).  Start that fill now on the side stream: it runs underneath the forward kernels, which
            # are not write-bound, and backward only has to scatter.
            # join the fill here (device-side wait, no host sync): from now on the buffer is ordinary main-stream
            # memory, so it needs no record_stream and may be dropped safely if backward never runs
            stream.wait_stream(st["side"])
            ctx.prefilled = [st["grad_buf"]]
        return loss.reshape(())

    @staticmethod
    def _forward_fused(ctx, rep, st, dims, bank, layout, dev, sp, p2p=None):
        """Production path: one allocation, one FFI call (arco_forward), no tensor views besides the loss."""
        Cn, Q, N, D = dims.classes, dims.queries, dims.negatives, dims.feat
        al = lambda n: (n + 255) & ~255
        o_proto = al(layout.total_bytes)
        o_ia = o_proto + al(Cn * (D + 1) * 8)
        o_in = o_ia + al(Cn * Q * 4)
        o_loss = o_in + al(Cn * Q * max(N, 1) * 4)
        o_pix = o_loss + 256
        o_ga = o_pix + al(Cn * Q * 4)
        o_mom = o_ga + al(Cn * Q * D * 4)
        mom = st["momentum"]
        total = o_mom + (al(Cn * Q * D * 4) + 256 if mom is not None else 0)
        buf = torch.empty(total, dtype=torch.uint8, device=dev)
        base = buf.data_ptr()
        io = _cabi.StepIO()
        io.rep, io.rep_teacher = st["rep_data"].data_ptr(), st["rep_teacher"].data_ptr()
        io.label_l = st["label_l"].data_ptr() if st["label_l"] is not None else None
        io.label_u = st["label_u"].data_ptr() if st["label_u"] is not None else None
        io.prob_l = st["prob_l"].data_ptr() if st["prob_l"] is not None else None
        io.prob_u = st["prob_u"].data_ptr() if st["prob_u"] is not None else None
        io.low_mask, io.high_mask = st["low_mask"].data_ptr(), st["high_mask"].data_ptr()
        io.proto_sums, io.idx_anchor, io.idx_neg = base + o_proto, base + o_ia, base + o_in
        io.loss, io.anchor_pix, io.grad_anchor = base + o_loss, base + o_pix, base + o_ga
        io.logits = None
        grad_buf = None
        if st["prefill"]:
            grad_buf = torch.empty(rep.shape, dtype=rep.dtype, device=dev)
            io.grad_prefill = grad_buf.data_ptr()
        if mom is not None:
            mom_on = (mom != 0).any().to(torch.int32).reshape(1)
            proto_out = buf[o_mom: o_mom + Cn * Q * D * 4].view(torch.float32).view(mom.shape)
            proto_out.zero_()
            io.momentum, io.momentum_on, io.proto_out = mom.data_ptr(), mom_on.data_ptr(), proto_out.data_ptr()
            # the reference forms (1 - ema_decay) as a Python double before it meets the float32 tensor (:491-495)
            io.ema_decay, io.ema_keep = float(st["ema_decay"]), float(1.0 - st["ema_decay"])
        io.seed, io.step = st["seed"], st["step"]
        io.delta_p, io.delta_n, io.temp = DELTA_P, float(st["delta_n"]), float(st["temp"])
        io.low_rank, io.high_rank, io.func = LOW_RANK, HIGH_RANK, st["func"]
        if p2p is not None:
            global P2P_EXCHANGE_USED
            P2P_EXCHANGE_USED = True
            p2p["seq"] += 1
            io.exchange_peers = p2p["peers"].data_ptr()
            io.exchange_local = p2p["buf"].data_ptr() + (p2p["seq"] & 1) * p2p["slot"] * 8
            io.exchange_seq, io.exchange_slot = p2p["seq"] | p2p.get("seq_flags", 0), p2p["slot"]
            io.exchange_rank, io.exchange_world = p2p["rank"], p2p["world"]
        _cabi.check(_cabi.lib.arco_forward(C.byref(dims), C.byref(io), C.byref(bank.c_struct), base, sp), "arco_forward")
        plan_view = buf[layout.plan: layout.plan + C.sizeof(_cabi.Plan)]
        bank.post_step(plan_view)
        st["plan_view"] = plan_view
        if mom is not None:
            n_valid = buf[layout.plan + 384: layout.plan + 388].view(torch.int32)      # arco_plan.n_valid
            st["prototype_out"] = torch.where(n_valid <= 1, mom, proto_out)
        ctx.fused = (buf, base + o_ga, base + o_pix)
        if st["debug"] is not None:
            # views into the step's packed buffer (parity tests / bench.py's cross-rank check of the fused path)
            st["debug"].update(
                ws=buf, layout=layout, dims=dims,
                proto_sums=buf[o_proto: o_proto + Cn * (D + 1) * 8].view(torch.float64).view(Cn, D + 1),
                idx_anchor=buf[o_ia: o_ia + Cn * Q * 4].view(torch.int32).view(Cn, Q),
                idx_neg=buf[o_in: o_in + Cn * Q * max(N, 1) * 4].view(torch.int32).view(Cn, Q * max(N, 1)),
                anchor_pix=buf[o_pix: o_pix + Cn * Q * 4].view(torch.int32).view(Cn, Q),
                grad_anchor=buf[o_ga: o_ga + Cn * Q * D * 4].view(torch.float32).view(Cn, Q, D), logits=None)
        ctx.dims = dims
        ctx.rep_shape = rep.shape
        ctx.rep_dtype = rep.dtype
        ctx.prefilled = [grad_buf] if grad_buf is not None else None
        ctx.sparse = st["sparse"]
        return buf[o_loss: o_loss + 4].view(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        if getattr(ctx, "fused", None) is not None:
            buf, ga_ptr, pix_ptr = ctx.fused
            dev = buf.device
        else:
            g_anchor, pix = ctx.saved_tensors
            dev = g_anchor.device
            ga_ptr, pix_ptr = g_anchor.data_ptr(), pix.data_ptr()
        go = grad_out.detach().to(torch.float32).contiguous()
        stream = torch.cuda.current_stream(dev)
        sp = stream.cuda_stream
        pre = ctx.prefilled
        if ctx.sparse is not None:
            # opt-in sparse-gradient contract: the op owns grad_rep across steps; it is zero except at the previous
            # step's anchor pixels, which are cleared before this step's scatter (no P*D*e-byte fill)
            grad_rep, prev_pix = ctx.sparse
            _cabi.check(_cabi.lib.arco_grad_scatter_sparse(C.byref(ctx.dims), ga_ptr, pix_ptr, go.data_ptr(),
                                                           grad_rep.data_ptr(), prev_pix.data_ptr(), sp),
                        "arco_grad_scatter_sparse")
        elif pre is not None and pre[0] is not None:
            grad_rep = pre[0]
            pre[0] = None                                   # a second backward (retain_graph) takes the slow path
            _cabi.check(_cabi.lib.arco_grad_scatter_add(C.byref(ctx.dims), ga_ptr, pix_ptr,
                                                        go.data_ptr(), grad_rep.data_ptr(), sp), "arco_grad_scatter_add")
        else:
            grad_rep = torch.empty(ctx.rep_shape, dtype=ctx.rep_dtype, device=dev)
            _cabi.check(_cabi.lib.arco_grad_scatter(C.byref(ctx.dims), ga_ptr, pix_ptr, go.data_ptr(),
                                                    grad_rep.data_ptr(), sp), "arco_grad_scatter")
        return grad_rep, None


def compute_contra_memobank_loss(
    rep,
    label_l,
    label_u,
    prob_l,
    prob_u,
    low_mask,
    high_mask,
    memobank,
    queue_prtlis,
    queue_size,
    rep_teacher,
    momentum_prototype=None,
    i_iter=0,
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
    """Stratified pixel/voxel contrastive loss with memory bank -- same contract as the reference.

    Arguments 1-18 are the reference's (loss_helper_3d.py:271-290).  ``label_l`` / ``label_u`` are the
    int64 one-hot maps ``[B_x, C, *S]`` the trainers pass, or -- cheaper, 8 B instead of 8*C B per pixel --
    plain integer label maps ``[B_x, *S]`` (ignore label -1 is folded into class 0 like the trainers'
    ``label_onehot``); in that case the number of classes is taken from ``prob_l``.
    Returns ``(new_keys, loss)`` (``loss.backward()`` yields a dense ``rep.grad``), or
    ``(prototype, new_keys, loss)`` when ``momentum_prototype`` ``[C, Q, 1, D]`` is given (EMA prototypes,
    loss_helper_3d.py:488-497,512-513; ``i_iter`` must then be >= 1 as in the reference).

    Keyword-only extensions: ``process_group`` (batch-sharded multi-GPU: one all-reduce of the per-class
    prototype sums), ``seed`` (Philox seed of the in-kernel sampler; defaults to torch's CUDA seed),
    ``_inject`` / ``_debug`` (parity tests), and ``sparse_grad``.

    ``sparse_grad=True`` is an OPT-IN deviation from the autograd contract that removes the largest byte term of the
    step (the P*D*e-byte zero fill of the dense ``grad_rep``): the gradient handed to
    autograd is a buffer the op keeps per memory bank and reuses every step -- all zero except the <= C*Q anchor
    pixels of the current step; the previous step's pixels are cleared first.  The VALUES are identical to the default
    path; what changes is ownership: the tensor is only valid until the next backward of this op on the same bank.
    That is exactly how the trainers use it (``rep`` is the non-leaf output of ``q_representation``,
    train_arco_2d.py:317-329: its gradient is consumed by the conv backward of the same ``loss.backward()``); do not
    use it when ``rep`` is a leaf whose ``.grad`` you keep or accumulate across steps.
    """
    if not (torch.is_tensor(rep) and rep.is_cuda):
        raise RuntimeError("arco_b200.compute_contra_memobank_loss needs CUDA tensors: there is no CPU fallback")
    if rep.dim() not in (4, 5):
        raise ValueError(f"rep must be [B,D,H,W] or [B,D,H,W,Z], got {tuple(rep.shape)}")
    if rep.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError(f"rep must be float32 or bfloat16, got {rep.dtype}")
    if rep_teacher.shape != rep.shape or rep_teacher.dtype != rep.dtype or rep_teacher.device != rep.device:
        raise ValueError("rep_teacher must match rep in shape, dtype and device")
    dev = rep.device
    B, D = rep.shape[0], rep.shape[1]
    spatial = tuple(rep.shape[2:])
    S = 1
    for s in spatial:
        S *= int(s)
    n_lab, n_unlab = int(label_l.shape[0]), int(label_u.shape[0])
    if n_lab + n_unlab != B:
        raise ValueError(f"label_l ({n_lab}) + label_u ({n_unlab}) images != rep batch ({B})")
    Cn = int(prob_l.shape[1]) if prob_l.numel() else int(prob_u.shape[1])
    if label_l.dim() == rep.dim():
        label_kind = _cabi.LABEL_ONEHOT_I64
        if tuple(label_l.shape[1:]) != (Cn,) + spatial or tuple(label_u.shape[1:]) != (Cn,) + spatial:
            raise ValueError("one-hot labels must be [B_x, C, *spatial] matching prob and rep")
    elif label_l.dim() == rep.dim() - 1:
        label_kind = _cabi.LABEL_INDEX_I64
        if tuple(label_l.shape[1:]) != spatial or tuple(label_u.shape[1:]) != spatial:
            raise ValueError("integer label maps must be [B_x, *spatial] matching rep")
    else:
        raise ValueError("label_l must be one-hot [B_l,C,*S] or an integer map [B_l,*S]")
    if label_l.dtype != torch.int64 or label_u.dtype != torch.int64:
        raise ValueError("labels must be int64 (the trainers pass label_onehot(...).long())")
    for name, t, n in (("prob_l", prob_l, n_lab), ("prob_u", prob_u, n_unlab)):
        if t.dtype != torch.float32 or tuple(t.shape) != (n, Cn) + spatial:
            raise ValueError(f"{name} must be float32 [{n},{Cn},*spatial], got {t.dtype} {tuple(t.shape)}")
    for name, t in (("low_mask", low_mask), ("high_mask", high_mask)):
        if t.dtype != torch.float32 or tuple(t.shape) != (B, 1) + spatial:
            raise ValueError(f"{name} must be float32 [B,1,*spatial], got {t.dtype} {tuple(t.shape)}")
    for t in (label_l, label_u, prob_l, prob_u, low_mask, high_mask):
        if t.device != dev:
            raise ValueError("all tensor arguments must live on rep's device")
    if Cn > _cabi.MAX_CLASSES:
        raise ValueError(f"at most {_cabi.MAX_CLASSES} classes are supported, got {Cn}")
    if D % 4 != 0 or D > 512:
        raise ValueError(f"feature size D must be a multiple of 4 and <= 512, got {D}")
    if B * S >= 2 ** 31:
        raise ValueError("more than 2^31 pixels per call")
    if num_queries <= 0 or num_negatives < 0 or temp <= 0:
        raise ValueError("num_queries must be > 0, num_negatives >= 0, temp > 0")
    if len(memobank) != Cn:
        raise ValueError(f"memobank has {len(memobank)} classes, prob has {Cn}")

    mom, ema_decay = None, 0.0
    if momentum_prototype is not None:
        if tuple(momentum_prototype.shape) != (Cn, int(num_queries), 1, D) or momentum_prototype.device != dev:
            raise ValueError(f"momentum_prototype must be [{Cn},{int(num_queries)},1,{D}] on rep's device")
        mom = momentum_prototype.detach().to(torch.float32).contiguous()
        # reference: min(1 - 1/i_iter, 0.999), evaluated only when the tensor is not all zero (ZeroDivisionError at i_iter=0)
        ema_decay = min(1.0 - 1.0 / i_iter, 0.999) if i_iter != 0 else float("nan")
    with torch.cuda.device(dev):
        bank = DeviceMemoryBank.adopt(memobank, queue_prtlis, queue_size, D, dev, rep.dtype)
        bank.poll()                         # mirror finished steps (non-blocking): queue_prtlis, label errors
        bank.begin_step()
        sampler_seed, sampler_step = _sampler_stream(seed, bank, dev)
        key = (n_lab, n_unlab, Cn, D, S, int(num_queries), int(num_negatives), rep.dtype, label_kind, dev.index)
        cached = _GEOMETRY.get(key)
        if cached is None:
            dims = _cabi.Dims(n_lab, n_unlab, Cn, D, S, int(num_queries), int(num_negatives),
                              _cabi.BF16 if rep.dtype == torch.bfloat16 else _cabi.F32, label_kind)
            cached = _GEOMETRY[key] = (dims, _cabi.workspace_layout(dims))
        dims, layout = cached
        rep_data = rep.detach().contiguous()
        state = dict(
            dims=dims, layout=layout, bank=bank,
            label_l=label_l.contiguous() if n_lab else None, label_u=label_u.contiguous() if n_unlab else None,
            prob_l=prob_l.contiguous() if n_lab else None, prob_u=prob_u.contiguous() if n_unlab else None,
            low_mask=low_mask.contiguous(), high_mask=high_mask.contiguous(),
            rep_teacher=rep_teacher.detach().contiguous(), rep_data=rep_data,
            delta_n=delta_n, temp=temp, func=_FUNC.get(func, _cabi.FUNC_UNIFORM),
            seed=sampler_seed, step=sampler_step,
            group=process_group, inject=_inject, debug=_debug, momentum=mom, ema_decay=ema_decay,
            # only worth its three extra host calls when the step is bandwidth- rather than launch-bound
            prefill=bool(rep.requires_grad and torch.is_grad_enabled() and _PREFILL_GRAD
                         and rep.numel() * rep.element_size() >= _PREFILL_MIN_BYTES),
        )
        state["sparse"] = _sparse_state(bank, rep, Cn * int(num_queries)) if (sparse_grad and rep.requires_grad and torch.is_grad_enabled()) else None
        if state["sparse"] is not None:
            state["prefill"] = False
        loss = _ContraLoss.apply(rep, state)
    keys = LazyKeys(bank, Cn, state["plan_view"])
    if mom is not None:
        return state["prototype_out"].to(momentum_prototype.dtype), keys, loss
    return keys, loss


def compute_contra_memobank_loss_from_logits(
    rep,
    train_l_label,
    train_u_aug_label,
    pred_l_teacher,
    pred_u_teacher,
    pred_u,
    alpha_t,
    memobank,
    queue_prtlis,
    queue_size,
    rep_teacher,
    delta_n=1.0,
    func="asmc",
    num_queries=256,
    num_negatives=512,
    temp=0.5,
    *,
    seed: Optional[int] = None,
    sparse_grad: bool = False,
    _debug: Optional[dict] = None,
):
    """The trainers' mask preparation AND the loss in one op, from the raw tensors (``train_arco_2d.py:345-398``):

        prob_*_teacher = softmax(pred_*_teacher); entropy = H(softmax(pred_u)); low / high thresholds = np.percentile(entropy[valid],
        alpha_t / 100 - alpha_t); low_mask_all / high_mask_all; label_onehot;
        compute_contra_memobank_loss(rep, label_l, label_u, prob_l_teacher, prob_u_teacher, low_mask_all, high_mask_all, ...)

    Equivalent -- bit for bit -- to ``prepare_contrast_inputs(...)`` followed by ``compute_contra_memobank_loss(...)`` on its
    outputs, but no probability, mask or one-hot tensor is written: the classify kernel reads the C teacher logits, the int64
    label (ignore label negative) and the student entropy per pixel and forms softmax and masks in registers
    (``arco_classify_plan_logits``).  ``pred_*`` are float32 ``[B_x, C, *S]`` logits (2 <= C <= 8), labels int64 ``[B_x, *S]``.
    Returns ``(new_keys, loss)``; ``_debug`` receives the thresholds."""
    from .prepare import _q32
    if not (torch.is_tensor(rep) and rep.is_cuda):
        raise RuntimeError("arco_b200 needs CUDA tensors: there is no CPU fallback")
    if rep.dim() not in (4, 5) or rep.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("rep must be float32 / bfloat16 [B,D,*S]")
    if rep_teacher.shape != rep.shape or rep_teacher.dtype != rep.dtype or rep_teacher.device != rep.device:
        raise ValueError("rep_teacher must match rep in shape, dtype and device")
    dev = rep.device
    B, D = int(rep.shape[0]), int(rep.shape[1])
    spatial = tuple(rep.shape[2:])
    S = 1
    for x in spatial:
        S *= int(x)
    n_lab, n_unlab = int(train_l_label.shape[0]), int(train_u_aug_label.shape[0])
    if n_lab + n_unlab != B:
        raise ValueError("labelled + unlabelled images != rep batch")
    Cn = int(pred_u_teacher.shape[1]) if n_unlab else int(pred_l_teacher.shape[1])
    if not 2 <= Cn <= 8:
        raise ValueError(f"the logits-in op supports 2 <= C <= 8 classes, got {Cn}; use prepare_contrast_inputs + compute_contra_memobank_loss")
    if S % 4 != 0:
        raise ValueError("the logits-in op needs prod(spatial) % 4 == 0")
    for name, t, n in (("pred_l_teacher", pred_l_teacher, n_lab), ("pred_u_teacher", pred_u_teacher, n_unlab), ("pred_u", pred_u, n_unlab)):
        if not (t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (n, Cn) + spatial):
            raise ValueError(f"{name} must be CUDA float32 [{n},{Cn},*spatial], got {t.dtype} {tuple(t.shape)}")
    for name, t, n in (("train_l_label", train_l_label, n_lab), ("train_u_aug_label", train_u_aug_label, n_unlab)):
        if t.dtype != torch.int64 or tuple(t.shape) != (n,) + spatial:
            raise ValueError(f"{name} must be int64 [{n},*spatial]")
    if D % 4 != 0 or D > 512 or len(memobank) != Cn or B * S >= 2 ** 31 or num_queries <= 0 or num_negatives < 0 or temp <= 0:
        raise ValueError("bad D / memobank / sizes (see compute_contra_memobank_loss)")
    with torch.cuda.device(dev):
        bank = DeviceMemoryBank.adopt(memobank, queue_prtlis, queue_size, D, dev, rep.dtype)
        bank.poll()
        bank.begin_step()
        sampler_seed, sampler_step = _sampler_stream(seed, bank, dev)
        label_kind = _cabi.LABEL_INDEX_I64
        key = (n_lab, n_unlab, Cn, D, S, int(num_queries), int(num_negatives), rep.dtype, label_kind, dev.index)
        cached = _GEOMETRY.get(key)
        if cached is None:
            dims = _cabi.Dims(n_lab, n_unlab, Cn, D, S, int(num_queries), int(num_negatives),
                              _cabi.BF16 if rep.dtype == torch.bfloat16 else _cabi.F32, label_kind)
            cached = _GEOMETRY[key] = (dims, _cabi.workspace_layout(dims))
        dims, layout = cached
        state = dict(
            dims=dims, layout=layout, bank=bank,
            label_l=train_l_label.to(dev).contiguous() if n_lab else None,
            label_u=train_u_aug_label.to(dev).contiguous() if n_unlab else None,
            prob_l=None, prob_u=None, low_mask=None, high_mask=None,
            logits=dict(pred_l_teacher=pred_l_teacher.detach().contiguous(), pred_u_teacher=pred_u_teacher.detach().contiguous(),
                        pred_u=pred_u.detach().contiguous(), q_low=_q32(alpha_t), q_high=_q32(100 - alpha_t)),
            rep_teacher=rep_teacher.detach().contiguous(), rep_data=rep.detach().contiguous(),
            delta_n=delta_n, temp=temp, func=_FUNC.get(func, _cabi.FUNC_UNIFORM), seed=sampler_seed, step=sampler_step,
            group=None, inject=None, debug=_debug, momentum=None, ema_decay=0.0,
            prefill=bool(rep.requires_grad and torch.is_grad_enabled() and _PREFILL_GRAD
                         and rep.numel() * rep.element_size() >= _PREFILL_MIN_BYTES),
        )
        state["sparse"] = _sparse_state(bank, rep, Cn * int(num_queries)) if (sparse_grad and rep.requires_grad and torch.is_grad_enabled()) else None
        if state["sparse"] is not None:
            state["prefill"] = False
        loss = _ContraLoss.apply(rep, state)
        if _debug is not None:
            _debug["thresholds"] = state.get("thresholds")
    return LazyKeys(bank, Cn, state["plan_view"]), loss
