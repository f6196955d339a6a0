"""Build ``rendiation_b200/librdn_rt.so`` (the C-ABI library of include/rdn_rt.h) with nvcc for sm_100a.

In-tree build so the ``.so`` travels with the repository snapshot to the GPU box.  No torch dependency:
the library links only the (static) CUDA runtime.

    python -m rendiation_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "librdn_rt.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["bvh_builder.cpp", "accel.cpp", "traverse.cu", "compact.cu", "raygen.cu", "probe.cu", "pick.cu", "build_device.cu", "sbt.cu", "wavefront.cu", "capi.cu"]
# -fmad=false + IEEE div/sqrt: device arithmetic must round exactly like the reference's CPU code (DESIGN.md "Exactness");
# -ffp-contract=off does the same for the host builder/flattener.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
              "-I", INCLUDE]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")] + [os.path.join(INCLUDE, "rdn_rt.h"), __file__]
    nvcc = _nvcc()
    extra = os.environ.get("RDN_EXTRA_NVCC_FLAGS", "").split()  # experiments only (e.g. -DRDN_DEBUG_STEPS)
    jobs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + headers):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-Xptxas", "-v", "-c", sp, "-o", op]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            logs = list(ex.map(run, jobs))
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
        if verbose:
            print("\n".join(logs))
    if jobs or force or _stale(LIB, objs):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
