#!/usr/bin/env python
"""profiles/rNN_kernel_shares.md: per-kernel share of one step in the ncu launch list (serialised, cold caches) next to the
live CUDA-event stage timing of the same command's bench line.  python profiles/kernel_shares.py r02"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_OF = [("classify", "classify_plan"), ("proto_", "proto_enqueue"), ("sample_", "sample"), ("infonce", "infonce"),
            ("fill_zero", "grad_scatter"), ("grad_scatter", "grad_scatter"), ("grad_unscatter", "grad_scatter")]


def launches(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("arco::", "")
        if "Functor" in k or "at::" in k:
            continue                                   # torch fills of the bench scaffolding (L2 flush, ones_like)
        agg.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")) / 1000.0)
    return agg


def main():
    tag = sys.argv[1]
    bench = json.loads(open(os.path.join(ROOT, "gpurun_out", f"{tag}_final_bench.json")).read().strip().splitlines()[-1])
    blocks = {"trainstep": bench, "cityscapes": bench["configs"]["cityscapes"], "acdc2d_loss": bench["configs"]["acdc2d_loss"],
              "la3d": bench["configs"]["la3d"]}
    out = [f"# {tag}: kernel shares of one step -- ncu launch list vs live CUDA-event stage timing\n",
           f"`profiles/{tag}_launches_<workload>.csv` (ncu `--metrics gpu__time_duration.sum --clock-control none`, serialised, cold caches) "
           f"against the `stages` block of the driver-format bench line (`profiles/{tag}_bench_lines.md`; CUDA events on the launching "
           "stream, each C-ABI stage timed alone with its events and launches enqueued behind a device-side delay, so the host's issue time is not in the interval).\n"]
    for wl, blk in blocks.items():
        agg = launches(os.path.join(ROOT, "profiles", f"{tag}_launches_{wl}.csv"))
        step = sum(sum(v) / len(v) for v in agg.values())
        st = blk["stages"]
        ssum = sum(v["ms"] for k, v in st.items() if not k.startswith("_")) * 1e3
        out.append(f"\n## {wl}: {blk['ms_per_step']*1e3:.0f} us per step (eager), "
                   f"{((blk.get('cuda_graph_replay') or {}).get('ms_per_step') or 0)*1e3:.0f} us as a CUDA-graph replay\n")
        out.append("| kernel | launches | ncu us / launch | share of the ncu step | bench stage | stage us | share of the stage sum |\n|---|---|---|---|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
            stage = next((s for pat, s in STAGE_OF if pat in k), "?")
            sms = st.get(stage, {}).get("ms", 0) * 1e3
            out.append(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {100*sum(v)/len(v)/step:.1f} % | {stage} | {sms:.1f} | {100*sms/ssum:.1f} % |")
        out.append(f"\nncu step (one launch of each kernel): {step:.0f} us; stage sum: {ssum:.0f} us.")
    with open(os.path.join(ROOT, "profiles", f"{tag}_kernel_shares.md"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
