"""In-tree build of libarco_b200.so (hand-written CUDA for sm_100a behind a C ABI).

``python -m arco_b200.build`` or ``arco_b200.build.build()``.  nvcc cross-compiles without a GPU;
the resulting ``arco_b200/lib/libarco_b200.so`` is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libarco_b200.so")
SOURCES = ["cabi.cu", "classify.cu", "scan_plan.cu", "proto_enqueue.cu", "proto_tc.cu", "proto_tc32.cu", "sampler.cu", "infonce.cu", "sim_dense.cu", "prepare.cu", "revisit.cu", "stepterms.cu", "producers.cu", "allreduce.cu", "grad.cu", "forward.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra=None, out: str = LIB) -> str:
    """``extra``: additional nvcc flags (e.g. ``-DARCO_TC_SUB=1``) for A/B builds written to ``out``."""
    os.makedirs(LIBDIR, exist_ok=True)
    extra = list(extra or [])
    tag = "" if out == LIB else "." + os.path.basename(out).replace(".so", "")
    headers = [os.path.join(CSRC, "arco_common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(CSRC, "proto_tail.cuh"),
               os.path.join(CSRC, "plan_common.cuh"),
               os.path.join(HERE, "..", "include", "arco_b200.h")]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", tag + ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or force or _stale(out, objs):
        cmd = [_nvcc(), "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
