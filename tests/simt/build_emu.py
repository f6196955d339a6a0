"""TEST INFRASTRUCTURE: build ``tests/simt/_build/librdn_rt_emu.so`` — the sources of rendiation_b200/csrc compiled by g++
with ``simt_emu.h`` force-included, so that the CUDA kernels run on the CPU (one fiber per CUDA thread) behind the same C-ABI.

Only ``tests/test_simt_emulation.py`` loads the result.  The package itself knows nothing about it and keeps failing loudly
without ``librdn_rt.so``.

    python tests/simt/build_emu.py [--force]

The one textual change made to the sources on the way: every ``kernel<<<grid, block[, smem[, stream]]>>>(args);`` becomes
``::simt::launch(grid, block, [&]() { kernel(args); });`` (g++ has no launch syntax), and ``__noinline__`` is spelled out.
Everything else is handled by the header and the ``RDN_SIMT_EMU`` stand-ins for PTX-only operations in the sources.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "rendiation_b200", "csrc")
INCLUDE = os.path.join(ROOT, "include")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "librdn_rt_emu.so")
CUDA_INCLUDE = os.environ.get("CUDA_INCLUDE", "/usr/local/cuda/include")

SOURCES = ["bvh_builder.cpp", "accel.cpp", "traverse.cu", "compact.cu", "raygen.cu", "probe.cu", "pick.cu", "build_device.cu", "sbt.cu", "wavefront.cu", "capi.cu"]
EMU_SOURCES = ["simt_engine.cpp", "simt_cudart.cpp", "simt_selftest.cpp"]
CXXFLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-DRDN_SIMT_EMU", "-Wno-attributes",
            "-Wno-unknown-pragmas", "-I", INCLUDE, "-I", CSRC, "-I", HERE, "-isystem", CUDA_INCLUDE, "-include",
            os.path.join(HERE, "simt_emu.h")]


def _split_top_level(text: str) -> list[str]:
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


_LAUNCH = re.compile(r"(?P<kernel>\b[A-Za-z_]\w*(?:<[^<>;]*>)?)\s*<<<(?P<cfg>.*?)>>>\s*\((?P<args>.*?)\)\s*;", re.S)


def rewrite_launches(text: str) -> tuple[str, int]:
    """kernel<<<grid, block, ...>>>(args);  ->  ::simt::launch(grid, block, [&]() { kernel(args); });"""
    count = 0

    def sub(m: re.Match) -> str:
        nonlocal count
        cfg = _split_top_level(m.group("cfg"))
        if len(cfg) < 2:
            raise RuntimeError(f"cannot parse launch configuration: {m.group(0)}")
        count += 1
        # keep the number of lines (compiler messages and debug info keep pointing at the .cu lines)
        newlines = "\n" * m.group(0).count("\n")
        args = " ".join(m.group("args").split())
        return f"::simt::launch(dim3({cfg[0]}), dim3({cfg[1]}), [&]() {{ {m.group('kernel')}({args}); }});{newlines}"

    return _LAUNCH.sub(sub, text), count


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, defines: tuple[str, ...] = ()) -> str:
    """defines: extra -D macros (e.g. ("RDN_REF_LEAF_MAX_COUNT=1",)); each set of defines gets its own build directory"""
    defines = tuple(defines) or tuple(d for d in os.environ.get("RDN_SIMT_DEFINES", "").split() if d)
    # RDN_SIMT_CXXFLAGS: extra compiler / linker flags, e.g. "-fsanitize=address -fno-omit-frame-pointer" (then run python with
    # LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0): every access of
    # the kernels to "device" memory is then checked against the bounds of its cudaMalloc — a memcheck without a GPU
    extra = tuple(os.environ.get("RDN_SIMT_CXXFLAGS", "").split())
    tag = "_".join(re.sub(r"\W+", "-", d) for d in defines + extra)
    OUT = os.path.join(HERE, "_build", tag) if tag else os.path.join(HERE, "_build")
    LIB = os.path.join(OUT, "librdn_rt_emu.so")
    os.makedirs(OUT, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    headers += [os.path.join(INCLUDE, "rdn_rt.h"), os.path.join(HERE, "simt_emu.h"), __file__]
    jobs, objs = [], []
    for src in SOURCES + EMU_SOURCES:
        sp = os.path.join(CSRC if src in SOURCES else HERE, src)
        op = os.path.join(OUT, src.rsplit(".", 1)[0] + ".o")
        objs.append(op)
        if not (force or _stale(op, [sp] + headers)):
            continue
        if src.endswith(".cu"):
            with open(sp) as f:
                text, n = rewrite_launches(f.read())
            # (libstdc++ spells the attribute __noinline__ itself, so it cannot be a macro)
            text = re.sub(r"\b__noinline__\b", "__attribute__((noinline))", text)
            if "<<<" in text:
                raise RuntimeError(f"{src}: a kernel launch was not rewritten")
            gen = os.path.join(OUT, src + ".cpp")
            with open(gen, "w") as f:
                f.write(f'#line 1 "{sp}"\n' + text)
            sp = gen
        jobs.append(["g++", *CXXFLAGS, *extra, *[f"-D{d}" for d in defines], "-c", sp, "-o", op])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr[-6000:])
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run(["g++", "-shared", *extra, "-o", LIB, *objs, *([] if any("sanitize" in f for f in extra) else ["-Wl,--no-undefined"]), "-Wl,-Bsymbolic", "-lpthread"])  # -Bsymbolic: our cuda* stand-ins, not a libcudart torch loaded
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
