#!/usr/bin/env python
"""Turn gpurun_out/ artefacts (bench JSON lines, .ncu-rep captures) into the tracked summaries under profiles/.

    python profiles/summarize.py r01 gpurun_out/prof_trainstep_v3.ncu-rep [more.ncu-rep ...]
"""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]


def ncu_rows(rep):
    # a .csv is the raw page already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): the reports themselves are
    # too large to bring back (gpurun merges at most 64 MiB)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for m in METRICS:
            if m in hdr:
                d[m] = r[hdr.index(m)] + " " + units[hdr.index(m)]
        res.append(d)
    return res


def main():
    tag = sys.argv[1]
    lines = [f"# {tag}: ncu `--set full --clock-control none` summaries (per launch, cold-ish caches, serialised)\n"]
    traffic = {}
    for rep in sys.argv[2:]:
        lines.append(f"\n## {os.path.basename(rep)}\n")
        seen = set()
        for d in ncu_rows(rep):
            if d["kernel"] in seen:
                continue
            seen.add(d["kernel"])
            lines.append(f"\n### `{d['kernel'][:100]}`\n")
            lines.append("| metric | value |\n|---|---|")
            for m in METRICS:
                if m in d:
                    lines.append(f"| {m} | {d[m]} |")
            # DRAM bytes of the prototype kernel per launch -> profiles/traffic.json (bench.py's roofline.traffic)
            base = os.path.basename(rep)
            wl = ("acdc2d_trainstep" if "trainstep" in base else "cityscapes+coherent_masks" if "city" in base and "coherent" in base else
                  "cityscapes" if "city" in base else
                  "la3d" if "la3d" in base else "acdc2d_loss" if "acdc" in base else None)
            if wl and "proto_" in d["kernel"] and "finalize" not in d["kernel"]:
                def num(m):
                    v, u = d.get(m, "0 byte").split()[:2]
                    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic.setdefault(wl, {})["proto_enqueue"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
                traffic[wl]["source"] = f"{base}: {d['kernel'][:60]} dram__bytes_read.sum + dram__bytes_write.sum"
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if traffic:
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        old = json.load(open(tp)) if os.path.exists(tp) else {}
        old.update(traffic)
        with open(tp, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)
            f.write("\n")
    # bench lines
    out = [f"# {tag}: bench.py lines brought back from the B200 box\n"]
    for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "bench_*.json" if tag == "r01" else f"{tag}_final_bench*.json"))):
        try:
            j = json.loads(open(path).read().strip().splitlines()[-1])
        except Exception:
            continue
        out.append(f"\n## {os.path.basename(path)}: {j['config']['workload']} x{j['n_gpus']} -> {j['value']:.1f} {j['unit']}, "
                   f"{j['ms_per_step']:.3f} ms/step\n")
        out.append("| stage | ms | algorithmic MB | GB/s | frac of measured HBM peak |\n|---|---|---|---|---|")
        for k, v in j.get("stages", {}).items():
            if k.startswith("_"):
                continue
            out.append(f"| {k} | {v['ms']:.4f} | {v['alg_bytes']/1e6:.1f} | {v['gbs']:.0f} | {v['frac_hbm']:.3f} |")
        out.append(f"\nmeasured counts: {j.get('stages', {}).get('_measured')}\n")
        for cname, cv in (j.get("configs") or {}).items():
            out.append(f"\n### configs[{cname}] x{j['n_gpus']}: {cv['ms_per_step']:.4f} ms/step, {cv['value']:.1f} {cv['unit']}"
                       f" (CUDA-graph replay: {(cv.get('cuda_graph_replay') or {}).get('ms_per_step')})\n")
            if cv.get("stages"):
                out.append("| stage | ms | algorithmic MB | GB/s | frac of measured HBM peak |\n|---|---|---|---|---|")
                for k, v in cv["stages"].items():
                    if not k.startswith("_"):
                        out.append(f"| {k} | {v['ms']:.4f} | {v['alg_bytes']/1e6:.1f} | {v['gbs']:.0f} | {v['frac_hbm']:.3f} |")
        out.append("```json\n" + json.dumps({k: j[k] for k in j if k not in ("stages", "configs")}) + "\n```")
    with open(os.path.join(ROOT, "profiles", f"{tag}_bench_lines.md"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
