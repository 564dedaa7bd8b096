#!/usr/bin/env python
"""profiles/rNN_sass_excerpts.md: Blackwell-specific SASS mnemonics per kernel of libarco_b200.so (cuobjdump -sass).
    python profiles/sass_excerpts.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "arco_b200", "lib", "libarco_b200.so")
WANT = re.compile(r"^(UTC\w*MMA|UTCBAR|UTCATOMSWS|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|SYNCS|HMMA|REDUX|MATCH|LDSM|UTCCP|STG\.E\.ENL2\.256)")
# one representative instantiation per kernel family (first match in the library)
FAMILIES = ["classify_small_kernel", "classify_kernel", "proto_tc_kernel", "proto_tc32_kernel", "keys_transform_kernelILb0", "keys_transform_kernelILb1",
            "sim_dense_kernel", "infonce_mma_kernel", "infonce_kernel", "infonce_lane_kernel", "revisit_dots_kernel", "sample_scan_kernel",
            "fill_zero_kernel"]


def main():
    tag = sys.argv[1]
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for ln in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m and cur:
            ins = m.group(1).strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            funcs[cur].append(ins)
    lines = [f"# {tag}: SASS evidence per tensor-core / TMA kernel (`cuobjdump -sass arco_b200/lib/libarco_b200.so`, sm_100a)\n",
             "Counts of the Blackwell-specific mnemonics per kernel, then the first occurrence of each in program order. `UTC*MMA` = `tcgen05.mma` "
             "(`UTCHMMA` kind::f16 / kind::tf32), `LDTM`/`STTM` = `tcgen05.ld`/`tcgen05.st`, `UTMALDG` = `cp.async.bulk.tensor` (TMA tile load), "
             "`UBLKCP` = `cp.async.bulk` (TMA 1-D), `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier, `HMMA` = `mma.sync` (legacy tensor path), "
             "`LDSM` = `ldmatrix`, `REDUX` = `redux.sync`, `STG.E.ENL2.256` = 256-bit store. Regenerate: `python profiles/sass_excerpts.py " + tag + "`.\n"]
    for fam in FAMILIES:
        name = next((f for f in funcs if fam in f), None)
        if not name:
            continue
        cnt, first = collections.Counter(), collections.OrderedDict()
        for ins in funcs[name]:
            m = WANT.match(ins)
            if m:
                key = ins.split()[0]
                key = re.sub(r"\.(64|128|32|TRANS64|x\d+|16816|F32|BF16|TF32|PHASECHK|ARRIVE|EXCH|TRYWAIT|A1T0|CCTL|NOINC|RED|ART0|1688|SUM|OR|ANY|U32|S32|M88|4|2|MT88).*", "", key)
                cnt[key] += 1
                first.setdefault(key, ins)
        lines.append(f"\n## `{fam.split('IL')[0]}` (`{name[:110]}`), {len(funcs[name])} instructions\n")
        if not cnt:
            lines.append("(none of the listed mnemonics)")
            continue
        lines.append("| mnemonic | count |\n|---|---|")
        for k, v in cnt.most_common():
            lines.append(f"| `{k}` | {v} |")
        lines.append("\n```")
        for k, ins in first.items():
            lines.append(ins)
        lines.append("```")
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpts.md"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
