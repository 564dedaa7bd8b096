#!/usr/bin/env python
"""Benchmark of the ARCO stratified contrastive loss hot path (fwd+bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

Contract (one JSON line on rank 0):
  metric/unit  BASELINE.json's metric: contrastive-loss fwd+bwd throughput in Mpixels/s (ms in ms_per_step)
  value        whole-job pixels / device time, inputs already resident in HBM
  e2e          same metric through the public op with HOST (pinned) buffers: H2D of every input and D2H of
               the loss inside the timed region
  roofline     dominant kernel (one-pass prototype reduce + key enqueue over rep_teacher) against the
               measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the oracle port (oracle/contra_oracle.py, torch CPU ops, all host threads) on a bounded sample
A "step" is one forward+backward of the loss over one synthetic batch.  Workloads: SURVEY.md section 8(d).
The default is BASELINE.json configs[1]'s shape (ACDC 2-D train step: batch 24 = 12 labelled + 12 unlabelled,
4 classes, 256x256, D=496 representation head, bf16 rep tensors); the U-Net stays in PyTorch and is not
part of the hot path.  Multi-GPU is weak scaling over the batch with one all-reduce of the prototype sums.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "contrastive_loss_fwd_bwd_throughput"
UNIT = "Mpixels/s"
KERNELS_PER_STEP = 7        # classify(+scan+plan), proto_enqueue(+finalize), sample_scan, sample_emit, infonce, fill_zero, grad_scatter


COLD_BANK = False


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="acdc2d_trainstep",
                    choices=["acdc2d_loss", "acdc2d_trainstep", "la3d", "cityscapes", "acdc2d_fullstep"])
    ap.add_argument("--no-fullstep", action="store_true",
                    help="skip the BASELINE config-2 block (full ARCO 2-D training step, PyTorch U-Net + this repo's loss ops vs the "
                         "reference's op mix; N=1 only)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--func", default="smc")
    ap.add_argument("--blocky", action="store_true", help="labels constant on 16-pixel tiles instead of iid")
    ap.add_argument("--coherent", action="store_true", help="low / high masks = bottom / top 20 %% of one smooth field per image (blobs) instead of iid pixels")
    ap.add_argument("--bank", default="full", choices=["full", "cold"],
                    help="full: banks pre-filled to capacity; cold: the trainers' initial one-row banks (reference-faithful "
                         "start, SURVEY.md section 8(d) config 3)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--no-aten-gpu", action="store_true",
                    help="skip timing the oracle port's torch/ATen ops ON THE GPU (the reference's own op mix, CPU banks "
                         "uploaded per class like loss_helper_3d.py:466): the like-for-like GPU baseline, on by default at N=1")
    ap.add_argument("--aten-gpu", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--cpu-budget", type=float, default=150.0,
                    help="--impl reference: seconds the whole CPU run may take; the batch is the largest n+n that fits")
    ap.add_argument("--no-configs", action="store_true",
                    help="only the headline workload: skip the `configs` block (acdc2d_loss, la3d, la3d cold bank, cityscapes)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason sampler: an NVML thread (1 ms period) that is started BEFORE the warm-up so that
    it is already sampling when the timed region begins; only samples taken between mark_begin() and mark_end()
    count.  Falls back to one nvidia-smi query if pynvml is unavailable."""

    def __init__(self, index):
        self.index = index
        self.samples = []          # (t, sm_mhz, reasons bitmask)
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.ready = False
        self.t0 = self.t1 = None

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop_flag:
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(r)))
                self.ready = True
                time.sleep(0.001)
        except Exception:
            self.ready = True

    def start(self):
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        t = time.perf_counter()
        while not self.ready and time.perf_counter() - t < 5.0:
            time.sleep(0.005)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= (self.t1 or 1e30)]
        if not inside:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(v) for v in out.strip().split(",")[:2]]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 0, "source": "nvidia-smi after the run (no NVML sample fell inside the timed region)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted(n for n, bit in bits.items() if any(s[2] & bit for s in inside))
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "source": "NVML thread, 1 ms period, samples inside the timed region only"}


# --------------------------------------------------------------------------------------------------
# reference / CPU arm: the oracle port on host cores, bounded sample of the same workload
# --------------------------------------------------------------------------------------------------
def _cpu_steps(workload, n_lab, n_unlab, steps, warmup, func):
    import torch

    import oracle
    from arco_b200.synth import bench_bank, bench_inputs
    spec, x = bench_inputs(workload, torch.device("cpu"), seed=1337, n_lab=n_lab, n_unlab=n_unlab)
    memobank, ptrs, caps = bench_bank(spec, cold=COLD_BANK)
    sampler = {"smc": oracle.grid_strata_sample, "asmc": oracle.grid_antithetic_sample}.get(func)
    rep = x["rep"].float().requires_grad_(True)           # CPU bf16 kernels are not what the reference ran on
    teacher = x["rep_teacher"].float()
    times = []
    for i in range(warmup + steps):
        rep.grad = None
        t0 = time.perf_counter()
        res = oracle.contra_memobank_loss(
            rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            memobank, ptrs, caps, teacher, delta_n=0.97, sampler=sampler, num_queries=spec.queries,
            num_negatives=spec.negatives, temp=0.5)
        res.loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return spec, sum(times) / len(times)


def run_cpu(workload, steps, warmup, func, budget_s=150.0):
    """The oracle port on all host threads.  The batch is the LARGEST n+n (labelled + unlabelled, up to the workload's
    own) whose `warmup + steps` passes fit `budget_s`, found by timing one 1+1 pass first (CPU cost is close to linear in
    the batch: per-class fixed work -- bank gather, samplers -- amortises, so a larger sample can only favour the CPU)."""
    import torch
    from arco_b200.synth import WORKLOADS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full = WORKLOADS[workload]
    t_probe0 = time.perf_counter()
    spec, t1 = _cpu_steps(workload, 1, 1, 1, 0, func)
    probe_s = time.perf_counter() - t_probe0
    n = 1
    if full["n_lab"] == full["n_unlab"]:
        left = budget_s - probe_s
        while n < full["n_lab"] and (n + 1) * t1 * (steps + warmup) * 1.15 <= left:
            n += 1
    if n > 1 or steps > 1 or warmup > 0:
        spec, t = _cpu_steps(workload, n, n, steps, warmup, func)
    else:
        t = t1
    ms = 1e3 * t
    px = spec.pixels
    same = (n == full["n_lab"] and n == full["n_unlab"])
    sample = (f"{workload} shape, batch {n} labelled + {n} unlabelled of the workload's {full['n_lab']}+{full['n_unlab']} "
              f"({px} pixels/step, D={spec.feat}, C={spec.classes}, Q=256, N=512, fp32 on CPU), "
              f"{steps} timed steps after {warmup} warm-up, oracle port (torch CPU ops); batch chosen to fit a {budget_s:.0f} s budget")
    return dict(value=px / (ms * 1e-3) / 1e6, unit=UNIT, cores=cores, kind="port", sample=sample, ms_per_step=ms,
                threads=torch.get_num_threads(), batch=[n, n],
                same_config=(True if same else f"batch reduced to {n}+{n} of {full['n_lab']}+{full['n_unlab']} (CPU time box); "
                             "everything else identical; fp32 on CPU where the GPU arm stores rep in " + full["dtype"])), spec


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from arco_b200.synth import WORKLOADS
    base, spec = run_cpu(args.workload, max(1, args.steps), max(0, args.warmup), args.func, budget_s=args.cpu_budget)
    full = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": sum(base["batch"]), "labelled_per_gpu": base["batch"][0],
                   "classes": spec.classes, "spatial": list(spec.spatial), "feat": spec.feat, "rep_storage": "f32",
                   "queries": spec.queries, "negatives": spec.negatives, "func": args.func,
                   "workload_batch": full["n_lab"] + full["n_unlab"], "same_config": base["same_config"],
                   "note": "CPU arm: oracle port of the reference loss (same ATen op mix) on host cores; "
                           "the Python reference itself cannot travel to the GPU box", "sample": base["sample"]},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def measure_workload(torch, dist, arco_b200, _cabi, name, dev, rank, world, group, steps, warmup, func, blocky, cold,
                     with_stages=True, check_ranks=False, sparse_grad=False, graph=True, index_labels=False, coherent=False):
    """One workload, device-timed: W warm-up steps, then `steps` forward+backward passes, each bracketed by CUDA events on
    the launching stream; max over ranks.  Returns the block that goes into the JSON line (headline or `configs`)."""
    from arco_b200.synth import bench_bank, bench_inputs
    spec, x = bench_inputs(name, dev, seed=1337 + rank, blocky=blocky, coherent=coherent)
    memobank, ptrs, caps = bench_bank(spec, seed=1337 + rank, cold=cold)
    rep = x["rep"].requires_grad_(True)
    P = spec.pixels
    if index_labels:
        # the op's cheaper label input: int64 class maps [B,*S] (8 B/pixel) instead of the trainers' int64 one-hot (8*C B/pixel)
        x["label_l"] = x["labels"][: spec.n_lab].clamp_min(0).contiguous()
        x["label_u"] = x["labels"][spec.n_lab:].clamp_min(0).contiguous()
    kw = dict(delta_n=0.97, func=func, num_queries=spec.queries, num_negatives=spec.negatives, temp=0.5,
              process_group=group, seed=1337)
    if sparse_grad:
        kw["sparse_grad"] = True

    class _Consumer(torch.autograd.Function):
        # stands in for the layer that produced `rep` (q_representation's conv, train_arco_2d.py:317-329): it receives
        # grad_rep and passes nothing on, so the sparse-gradient variant is timed the way a trainer uses it (rep is NOT a
        # leaf; a leaf's AccumulateGrad would deep-copy the op-owned buffer)
        @staticmethod
        def forward(ctx, t):
            return t.view_as(t)

        @staticmethod
        def backward(ctx, g):
            return None

    def step(inputs=x, rep_t=rep):
        rep_t.grad = None
        rep_in = _Consumer.apply(rep_t) if sparse_grad else rep_t
        _, loss = arco_b200.compute_contra_memobank_loss(
            rep_in, inputs["label_l"], inputs["label_u"], inputs["prob_l"], inputs["prob_u"], inputs["low_mask"],
            inputs["high_mask"], memobank, ptrs, caps, inputs["rep_teacher"], **kw)
        loss.backward()
        return loss

    in_bytes = sum(x[k].numel() * x[k].element_size() for k in
                   ("rep", "rep_teacher", "label_l", "label_u", "prob_l", "prob_u", "low_mask", "high_mask"))
    flush = None
    l2_note = "inputs (%.0f MB) exceed the 126 MB L2" % (in_bytes / 1e6)
    if in_bytes < 400e6:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        l2_note = "L2 flushed (256 MB write) between timed steps"

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def replay_stats():
        import ctypes
        buf = (ctypes.c_int64 * 3)()
        _cabi.check(_cabi.lib.arco_forward_replay_stats(buf), "arco_forward_replay_stats")
        return list(buf)

    for _ in range(max(3, warmup)):
        step()
    sync_all()
    rs0 = replay_stats()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    sync_all()
    t_begin = time.perf_counter()
    for i in range(steps):
        if flush is not None:
            flush.fill_(i & 0xff)
        starts[i].record()
        step()
        ends[i].record()
    sync_all()
    t_end = time.perf_counter()
    rs1 = replay_stats()
    per_step = sorted(s.elapsed_time(e) for s, e in zip(starts, ends))
    total_ms = sum(per_step)
    t = torch.tensor([total_ms, per_step[len(per_step) // 2]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, median_ms = float(t[0].item()), float(t[1].item())
    out = {
        "workload": name + ("+cold_bank" if cold else "") + ("+blocky" if blocky else "") + ("+coherent_masks" if coherent else "") + ("+sparse_grad" if sparse_grad else "") + ("+index_labels" if index_labels else ""),
        "ms_per_step": total_ms / steps, "value": world * P * steps / (total_ms * 1e-3) / 1e6, "unit": UNIT, "steps": steps,
        # median of the per-step event times (max over ranks): the mean above is what the contract asks for, but with several
        # eager processes a short step (0.2 ms) picks up host-launch outliers of 0.5 ms+
        "ms_per_step_median": median_ms,
        "pixels_per_gpu": P, "rep_storage": spec.dtype, "l2": l2_note,
        # how arco_forward issued the timed steps' launches (rank 0): replayed as a captured CUDA graph (small shapes, see
        # DESIGN.md "Replay cache"), captured on the second sighting of a parameter tuple, or launched one by one
        "forward_issue": dict(zip(("graph_replays", "graphs_captured", "direct"), (a - b for a, b in zip(rs1, rs0)))),
    }
    if graph and world == 1:
        # The same public call, captured ONCE with torch.cuda.graph (forward + backward) and replayed: no launch parameter
        # changes between steps (the Philox stream and the host-mirror slot come from the bank's device step counter), so
        # this is the step a trainer that graph-captures its iteration runs.  Removes ~0.15 ms of per-step host work.
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
            for _ in range(3):
                g.replay()
            sync_all()
            gs = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
            ge = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
            for i in range(steps):
                if flush is not None:
                    flush.fill_(i & 0xff)
                gs[i].record()
                g.replay()
                ge[i].record()
            sync_all()
            gms = sum(a.elapsed_time(b) for a, b in zip(gs, ge)) / steps
            out["cuda_graph_replay"] = {"ms_per_step": gms, "value": P / (gms * 1e-3) / 1e6, "unit": UNIT,
                                        "what": "torch.cuda.graph capture of the public call (forward + backward), replayed; "
                                                "CUDA events per replay"}
            del g
        except Exception as e:                                  # noqa: BLE001 -- reported, never fatal for the bench line
            out["cuda_graph_replay"] = {"error": repr(e)[:300]}
    ctx = dict(spec=spec, x=x, rep=rep, memobank=memobank, ptrs=ptrs, caps=caps, kw=kw, flush=flush, sync_all=sync_all,
               window=(t_begin, t_end))
    if check_ranks and world > 1:
        # multi-GPU correctness bit the driver can see: one more step through the SAME fused path with the step's
        # global prototype sums kept, all-gathered, and compared bit for bit with rank 0's
        dbg = {"fused": True}
        rep.grad = None
        _, loss = arco_b200.compute_contra_memobank_loss(
            rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], memobank, ptrs, caps,
            x["rep_teacher"], _debug=dbg, **kw)
        mine = dbg["proto_sums"].clone().view(torch.int64).flatten()
        allp = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        same = all(bool(torch.equal(allp[0], a)) for a in allp)
        counts = dbg["proto_sums"][:, -1].sum().item()
        out["multi_gpu_check"] = {"proto_sums_bit_identical_on_all_ranks": same, "ranks": world,
                                  "global_low_valid_pixels": counts, "exchange": __import__("arco_b200.contra", fromlist=["x"]).EXCHANGE_PLANE}
        assert same, "global prototype sums differ between ranks"
    if with_stages and rank == 0:
        stages, roof = stage_timing(torch, _cabi, arco_b200, spec, x, rep, memobank, ptrs, caps, dev, flush,
                                    traffic_key=spec.name + ("+coherent_masks" if coherent else ""))
        out["stages"], out["roofline"] = stages, roof
        m = stages["_measured"]
        e_t = 2 if spec.dtype == "bf16" else 4
        whole = sum(v["alg_bytes"] for k, v in stages.items() if not k.startswith("_"))
        out["step_alg_bytes"] = whole
        out["step_frac_hbm"] = whole / (out["ms_per_step"] * 1e-3) / 1e9 / peaks()[0]
    return out, ctx


def main_fullstep(args):
    import torch
    assert torch.cuda.is_available(), "bench.py needs a GPU; there is no CPU fallback"
    blk = run_fullstep(max(3, args.steps))
    o = blk.get("arco_b200", {})
    print(json.dumps({
        "metric": "arco_2d_train_step_throughput", "value": o.get("value_mpixels_per_s"), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": 2, "ms_per_step": o.get("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": {"workload": "acdc2d_fullstep", "batch_per_gpu": 24, "classes": 4,
                                                            "spatial": [256, 256], "feat": 496}, "acdc2d_fullstep": blk}))


def main_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun exactly as the driver would
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU; there is no CPU fallback"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's version banner (printed at the VERSION and WARN debug levels) and any
        # other NCCL log text go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]               # the banner ignores NCCL_DEBUG_FILE; unset prints nothing (verified)
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    import gc

    import arco_b200
    from arco_b200 import _cabi

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    clocks.mark_begin()          # narrowed to the headline's timed region below
    head, ctx = measure_workload(torch, dist, arco_b200, _cabi, args.workload, dev, rank, world, group, args.steps,
                                 args.warmup, args.func, args.blocky, COLD_BANK, check_ranks=True, coherent=args.coherent)
    clocks.t0, clocks.t1 = ctx["window"]
    clk = clocks.stop() if rank == 0 else None
    spec, x, P = ctx["spec"], ctx["x"], ctx["spec"].pixels
    ms_per_step, value = head["ms_per_step"], head["value"]
    stages, roof = head.get("stages"), head.get("roofline")

    # ------------------------------------------------------------------ end-to-end with host buffers
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, arco_b200, spec, x, ctx["memobank"], ctx["ptrs"], ctx["caps"], dev, ctx["kw"], world,
                      args.e2e_steps or max(3, args.steps // 4), ctx["sync_all"])
        if world > 1:
            tt = torch.tensor([e2e["_total_ms"]], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e["_total_ms"] = float(tt.item())
        e2e["value"] = world * P * e2e["steps"] / (e2e.pop("_total_ms") * 1e-3) / 1e6

    aten = None
    if rank == 0 and world == 1 and not args.no_aten_gpu:
        aten = run_aten_gpu(torch, spec, x, dev, args.func)
        aten["speedup_of_this_op"] = aten["ms_per_step"] / ms_per_step
    bank_dtype = ctx["memobank"][0].bank.row_dtype
    del ctx, x
    gc.collect()
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ the other BASELINE.json configs, same run, same N
    configs = None
    if not args.no_configs:
        configs = {}
        todo = [("acdc2d_loss", False, False, False, False), ("la3d", False, False, False, False), ("la3d", True, False, False, False),
                ("cityscapes", False, False, False, False), (args.workload, COLD_BANK, True, False, False),
                (args.workload, COLD_BANK, False, True, False),
                # spatially coherent entropy masks (blobs, as on real predictions): the prototype pass skips the 32- / 64-pixel
                # steps that hold no low-valid or key pixel
                (args.workload, COLD_BANK, False, False, True), ("cityscapes", False, False, False, True)]
        for name, cold, sparse, idxlab, coh in todo:
            if name == args.workload and cold == COLD_BANK and not sparse and not idxlab and coh == args.coherent:
                continue
            blk, c2 = measure_workload(torch, dist, arco_b200, _cabi, name, dev, rank, world, group, max(5, args.steps // 2),
                                       args.warmup, "asmc" if name == "la3d" else args.func, args.blocky, cold,
                                       check_ranks=(name == "cityscapes" and not coh), sparse_grad=sparse, index_labels=idxlab,
                                       with_stages=not idxlab, coherent=coh or args.coherent)
            configs[blk["workload"]] = blk
            del c2
            gc.collect()
            torch.cuda.empty_cache()

    fullstep = None
    if rank == 0 and world == 1 and not args.no_fullstep and not args.no_configs:
        fullstep = run_fullstep(max(3, args.steps // 4))

    producers_blk = None
    if rank == 0 and world == 1 and not args.no_configs and args.workload == "acdc2d_trainstep":
        # the loss INCLUDING its producers (fea4 / q_representation), reference composition vs fused
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        try:
            import bench_producers
            producers_blk = bench_producers.run(args.workload)
        except Exception as e:                                  # noqa: BLE001 -- reported, never fatal for the bench line
            producers_blk = {"error": repr(e)[:300]}
        gc.collect()
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = run_cpu(args.workload, 2, 1, args.func, budget_s=25.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "same_config")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": args.workload, "batch_per_gpu": spec.batch, "labelled_per_gpu": spec.n_lab,
                "classes": spec.classes, "spatial": list(spec.spatial), "feat": spec.feat, "rep_storage": spec.dtype,
                "queries": spec.queries, "negatives": spec.negatives, "func": args.func,
                "labels": "blocky16" if args.blocky else "iid", "masks": "coherent blobs" if args.coherent else "iid Bernoulli(0.2)", "banks": ("cold: the trainers' initial one-row banks" if COLD_BANK else "pre-filled to capacity (50000/30000 rows)") + (", bf16-exact rows in a bf16 ring" if bank_dtype == torch.bfloat16 else ", fp32 ring"),
                "pixels_per_gpu": P, "parallelism": (f"batch-shard x{world}, 1 exchange of C*(D+1) fp64 (" + ("own exchange block inside the InfoNCE launch, over NVLink peer memory" if __import__("arco_b200.contra", fromlist=["x"]).P2P_EXCHANGE_USED else "NCCL all-reduce") + ")") if world > 1 else "single GPU",
                "l2": head["l2"], "timing": "CUDA events per step on the launching stream, max over ranks",
            },
            "clocks": clk, "gpu_launches": args.steps * (KERNELS_PER_STEP + (3 if world > 1 else 0)),   # N>1: + sample_scan / sample_emit / InfoNCE again, gated on a changed plan (early exit)
            "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "aten_gpu_baseline": aten, "stages": stages,
            "step_alg_bytes": head.get("step_alg_bytes"), "step_frac_hbm": head.get("step_frac_hbm"),
            "multi_gpu_check": head.get("multi_gpu_check"), "cuda_graph_replay": head.get("cuda_graph_replay"),
            "forward_issue": head.get("forward_issue"), "configs": configs,
            "acdc2d_fullstep": fullstep, "producers": producers_blk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_fullstep(steps):
    """BASELINE.json config 2 as stated: one synthetic ARCO 2-D training step (U-Net student + EMA teacher, FeatureExtractor,
    q_representation, CE + Dice, and all semi-supervised terms), bf16 autocast, batch 24, with the loss terms (a) as the
    reference composes them from ATen ops on the same GPU and (b) through this repository's CUDA ops (scripts/fullstep.py)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    try:
        import fullstep
        ours = fullstep.run("arco_b200", steps=steps, warmup=2)
        ref = fullstep.run("reference_ops", steps=max(2, steps // 2), warmup=1)
        return {"what": "train_arco_2d.py:284-431 on synthetic tensors: batch 12+12, 256x256, 4 classes, D=496, Q=256, N=512, K=36, "
                        "bf16 autocast backbones in PyTorch in both arms; wall clock with synchronize",
                "arco_b200": ours, "reference_ops_on_gpu": ref, "step_speedup": ref["ms_per_step"] / ours["ms_per_step"],
                "loss_terms_speedup": ref["ms_semi_supervised_terms"] / ours["ms_semi_supervised_terms"]}
    except Exception as e:                                      # noqa: BLE001 -- reported, never fatal for the bench line
        import traceback
        return {"error": repr(e)[:300], "trace": traceback.format_exc()[-600:]}


def stage_timing(torch, _cabi, arco_b200, spec, x, rep, memobank, ptrs, caps, dev, flush, iters=10, traffic_key=None):
    """Time every C-ABI stage on its own (CUDA events on the launching stream) and build the roofline
    entry of the dominant kernel from the algorithmic bytes of SURVEY.md section 8(d)."""
    C = ctypes
    lib = _cabi.lib
    bank = memobank[0].bank
    bank.settle()
    rep_dtype = _cabi.BF16 if spec.dtype == "bf16" else _cabi.F32
    S = 1
    for s in spec.spatial:
        S *= s
    dims = _cabi.Dims(spec.n_lab, spec.n_unlab, spec.classes, spec.feat, S, spec.queries, spec.negatives, rep_dtype, 0)
    L = _cabi.workspace_layout(dims)
    ws = torch.empty(L.total_bytes, dtype=torch.uint8, device=dev)
    Cn, Q, N, D = spec.classes, spec.queries, spec.negatives, spec.feat
    proto = torch.empty((Cn, D + 1), dtype=torch.float64, device=dev)
    idx_a = torch.empty((Cn, Q), dtype=torch.int32, device=dev)
    idx_n = torch.empty((Cn, Q * N), dtype=torch.int32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    g_anchor = torch.empty((Cn, Q, D), dtype=torch.float32, device=dev)
    pix = torch.empty((Cn, Q), dtype=torch.int32, device=dev)
    grad = torch.empty_like(x["rep"])
    go = torch.ones(1, dtype=torch.float32, device=dev)
    sp = torch.cuda.current_stream(dev).cuda_stream
    d, b = C.byref(dims), C.byref(bank.c_struct)
    fl = lambda t: t.contiguous().view(t.shape[0], t.shape[1], -1)
    ll, lu, pl, pu = fl(x["label_l"]), fl(x["label_u"]), fl(x["prob_l"]), fl(x["prob_u"])
    lm, hm = fl(x["low_mask"]), fl(x["high_mask"])
    rt, rs = x["rep_teacher"].contiguous(), x["rep"].detach().contiguous()
    calls = [
        ("classify_plan", lambda: lib.arco_classify_plan(d, ll.data_ptr(), lu.data_ptr(), pl.data_ptr(), pu.data_ptr(),
                                                         lm.data_ptr(), hm.data_ptr(), 0.3, 0.97, 3, 20, b, ws.data_ptr(), sp)),
        ("proto_enqueue", lambda: lib.arco_proto_enqueue(d, rt.data_ptr(), b, proto.data_ptr(), ws.data_ptr(), sp)),
        ("sample", lambda: lib.arco_sample(d, _cabi.FUNC_SMC, 1337, 7, idx_a.data_ptr(), idx_n.data_ptr(), ws.data_ptr(), sp)),
        ("infonce", lambda: lib.arco_infonce(d, rs.data_ptr(), b, proto.data_ptr(), idx_a.data_ptr(), idx_n.data_ptr(), 0.5,
                                             loss.data_ptr(), g_anchor.data_ptr(), pix.data_ptr(), None, ws.data_ptr(), sp)),
        ("grad_scatter", lambda: lib.arco_grad_scatter(d, g_anchor.data_ptr(), pix.data_ptr(), go.data_ptr(), grad.data_ptr(), sp)),
    ]
    acc = {name: 0.0 for name, _ in calls}
    # A stage is timed as the device sees it: start event, launch(es) and stop event are all ENQUEUED behind a ~100 us
    # device-side delay, so the interval holds the kernel(s) and their launch gaps but not the host's time to issue the call
    # (ctypes + cuTensorMapEncode + cudaLaunch, 10-20 us -- as much as a small kernel runs).  ncu's gpu__time_duration of the
    # same kernels (profiles/rNN_kernel_shares.md) is the cross-check.
    delay_cycles = int(100e-6 * 1.9e9)
    for it in range(iters + 2):
        for name, fn in calls:
            if flush is not None:
                flush.fill_(it & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(delay_cycles)
            e0.record()
            rc = fn()
            e1.record()
            _cabi.check(rc, name)
            e1.synchronize()
            if it >= 2:
                acc[name] += e0.elapsed_time(e1)
    ms = {k: v / iters for k, v in acc.items()}
    plan = _cabi.Plan.from_buffer_copy(ws[L.plan: L.plan + C.sizeof(_cabi.Plan)].cpu().numpy().tobytes())
    P = spec.pixels
    e_t = 2 if spec.dtype == "bf16" else 4
    e_bank = 2 if bank.row_dtype == torch.bfloat16 else 4
    P_lv = sum(int(plan.lv_count[c]) for c in range(Cn))
    K = sum(min(int(plan.n_key[c]), caps[c]) for c in range(Cn))
    Cv = sum(1 for j in range(Cn) if plan.slot_active[j])
    alg = {
        "classify_plan": P * (8 * Cn + 4 * Cn + 8) + P + 2 * Cn * L.n_tiles * 8,
        "proto_enqueue": P_lv * D * e_t + K * D * (e_t + e_bank) + P,
        "sample": Cv * (Q + Q * N) * 4,
        "infonce": Cv * Q * D * e_t + Cv * Q * N * D * e_bank + Cv * Q * D * 4,
        "grad_scatter": P * D * e_t + Cv * Q * D * (4 + 2 * e_t),
    }
    peak, peak_src = peaks()
    stages = {k: {"ms": ms[k], "alg_bytes": alg[k], "gbs": alg[k] / (ms[k] * 1e-3) / 1e9,
                  "frac_hbm": alg[k] / (ms[k] * 1e-3) / 1e9 / peak} for k in ms}
    stages["_measured"] = {"P_lv": P_lv, "K": K, "C_v": Cv, "sum_ms": sum(ms.values())}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get(traffic_key or spec.name, {}).get("proto_enqueue")
        except Exception:
            traffic = None
    k = "proto_enqueue"
    kname = ("arco::proto_tc_kernel (tcgen05 + TMA)" if spec.dtype == "bf16" and spec.classes <= 16 and spec.feat >= 64
             else "arco::proto_small_kernel" if spec.classes <= 3 and spec.feat in (16, 32)
             else "arco::proto_tc32_kernel (tcgen05 kind::tf32, hi + lo passes, TMA)"
             if spec.dtype != "bf16" and spec.feat > 128 and os.environ.get("ARCO_PROTO_TC32", "1") != "0"
             else "arco::proto_pipe_kernel")
    roof = {"kernel": kname + " (fp64 finalize in its tail)", "bound": "hbm",
            "achieved": stages[k]["gbs"], "peak": peak, "unit": "GB/s", "frac": stages[k]["frac_hbm"],
            "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_launch": alg[k], "ms_per_launch": ms[k],
            "bytes_formula": "P_lv*D*e_t + K*D*(e_t+e_bank) + P  (SURVEY.md section 8(d) teacher-read and key terms + 1 code byte per pixel)",
            "note": "rep_teacher is channel-first, so every 32-byte sector that holds one low-valid pixel must be fetched: "
                    "with the iid 20% masks of this workload that is ALL of P*D*e_t (coherent masks: steps without a needed pixel "
                    "are skipped); 'traffic' is the ncu-measured DRAM bytes of this kernel on this workload (profiles/traffic.json, "
                    "recorded, not live)"}
    if traffic:
        # the same launch against the DRAM bytes ncu counted for it (what the kernel really moved), next to the contract's
        # algorithmic figure
        roof["achieved_traffic"] = traffic / (ms[k] * 1e-3) / 1e9
        roof["frac_traffic"] = roof["achieved_traffic"] / peak
    return stages, roof


def run_aten_gpu(torch, spec, x, dev, func, steps=3):
    """The oracle port (same ATen op mix as the reference) with CUDA tensors: what running the reference's PyTorch loss on
    this B200 costs, including its per-class CPU->GPU bank upload and CPU samplers.  Reported, not optimised."""
    import oracle
    from arco_b200.synth import bench_bank
    memobank, ptrs, caps = bench_bank(spec, seed=4321, cold=COLD_BANK)
    sampler = {"smc": oracle.grid_strata_sample, "asmc": oracle.grid_antithetic_sample}.get(func)
    rep = x["rep"].detach().clone().requires_grad_(True)
    times = []
    for i in range(steps + 1):
        rep.grad = None
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res = oracle.contra_memobank_loss(
            rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], memobank, ptrs, caps,
            x["rep_teacher"], delta_n=0.97, sampler=sampler, num_queries=spec.queries, num_negatives=spec.negatives, temp=0.5)
        res.loss.backward()
        torch.cuda.synchronize(dev)
        if i > 0:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return {"value": spec.pixels / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "what": "oracle port (torch/ATen ops of the reference) on the same B200, full workload, wall clock with sync"}


def run_e2e(torch, arco_b200, spec, x, memobank, ptrs, caps, dev, kw, world, steps, sync_all):
    names = ("rep", "rep_teacher", "label_l", "label_u", "prob_l", "prob_u", "low_mask", "high_mask")
    host = {k: torch.empty(x[k].shape, dtype=x[k].dtype, pin_memory=True).copy_(x[k].detach()) for k in names}
    stage = {k: torch.empty_like(x[k].detach()) for k in names}
    host_loss = torch.empty(1, dtype=torch.float32, pin_memory=True)
    h2d = sum(host[k].numel() * host[k].element_size() for k in names)

    def one():
        for k in names:
            stage[k].copy_(host[k], non_blocking=True)
        rep = stage["rep"].requires_grad_(True)
        rep.grad = None
        _, loss = arco_b200.compute_contra_memobank_loss(
            rep, stage["label_l"], stage["label_u"], stage["prob_l"], stage["prob_u"], stage["low_mask"],
            stage["high_mask"], memobank, ptrs, caps, stage["rep_teacher"], **kw)
        loss.backward()
        host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        stage["rep"] = rep.detach()

    one()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    sync_all()
    return {"value": None, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": steps,
            "_total_ms": e0.elapsed_time(e1),
            "note": "public op arco_b200.compute_contra_memobank_loss; every input copied from pinned host memory "
                    "each step, loss read back; grad_rep stays on the device as in training"}


if __name__ == "__main__":
    a = parse()
    COLD_BANK = a.bank == "cold"
    if a.workload == "acdc2d_fullstep":
        main_fullstep(a)
    elif a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
