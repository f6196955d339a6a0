"""Issue-slot model of k_trace_ordered_rounds from the CPU emulation of the kernel (tests/simt) — no GPU needed.

    python tools/issue_model.py [c1|c2|c2s|c2t|c4] [--variant N] [--dump REGION] [--json]

static  : SASS instructions per marked source region of the kernel instantiation (nvdisasm line info of csrc/_obj/traverse.o,
          each instruction attributed to the kernel-body line it was inlined into, regions = the RDN_COST markers)
dynamic : how often a warp ISSUES each region when the emulated kernel runs the configuration (max over the lanes between two
          warp collectives, summed — simt_engine.cpp cost_flush) and how many lanes passed it
model   : sum(static x issues) = predicted warp-level instructions; lanes / issues = predicted active lanes per instruction.
The kernel is issue bound, so this is the quantity a kernel change has to lower.  Calibration: config 2 measured
by ncu at 143.8 M warp instructions and 19.1 active lanes (profiles/ncu_r1f_sass_mix_k_trace_ordered_c2.json).

What the model leaves out: instructions whose execution depends on data inside a region (early exits of the triangle test, the
slow path of the IEEE division), replays, and the placement of tiles on warps when several refill attempts interleave.
"""
from __future__ import annotations

import ctypes
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
SRC = os.path.join(ROOT, "rendiation_b200", "csrc", "traverse.cu")
INC = os.path.join(ROOT, "rendiation_b200", "csrc", "ordered_rounds.inc")
INC_AT = 0  # line of traverse.cu whose #include the loop's instructions are attributed to (set by region_lines)
OBJ = os.path.join(ROOT, "rendiation_b200", "csrc", "_obj", "traverse.o")


def region_names() -> list[str]:
    text = open(SRC).read()
    body = text[text.index("enum CostRegion {"):]
    body = body[:body.index("}")]
    return [n for n in re.findall(r"\b(COST_[A-Z0-9_]+)\b", body) if n != "COST_REGION_COUNT"]


def region_lines(names: list[str]):
    """markers of k_trace_ordered_rounds [(line, region)] in source order, and those inside triangle_test (an inlined callee)"""
    lines = open(SRC).read().split("\n")

    def collect(first, stop):
        out = []
        for i in range(first, len(lines)):
            m = re.search(r"RDN_COST\((COST_[A-Z0-9_]+)\)", lines[i])
            if m:
                out.append((i + 1, names.index(m.group(1))))
            if stop(lines[i], i):
                break
        return out

    k0 = next(i for i, l in enumerate(lines) if "k_trace_ordered_rounds(const __grid_constant__" in l)
    kernel = collect(k0, lambda l, i: l.startswith("}  // namespace"))
    # the round loop lives in ordered_rounds.inc, included once per loop of the kernel: its lines get positions right behind the
    # include line of the plain instantiations (the last one), include_line + line / 1e5
    global INC_AT
    INC_AT = max(i + 1 for i in range(k0, len(lines)) if lines[i].strip() == '#include "ordered_rounds.inc"')
    for i, l in enumerate(open(INC).read().split("\n")):
        m = re.search(r"RDN_COST\((COST_[A-Z0-9_]+)\)", l)
        if m:
            kernel.append((INC_AT + (i + 1) * 1e-5, names.index(m.group(1))))
    kernel.sort()
    t0 = next(i for i, l in enumerate(lines) if "bool triangle_test(" in l)
    t1 = next(i for i in range(t0, len(lines)) if lines[i] == "}")
    callee = collect(t0, lambda l, i: i >= t1)
    return kernel, callee, (t0 + 1, t1 + 1)


DUMP: dict[int, list[str]] = {}   # region -> its SASS lines (filled by static_counts; printed with --dump REGION)


def static_counts(template_args: str, names: list[str]) -> tuple[list[float], int]:
    """SASS instructions per region PER PASS of its marker, for the instantiation whose mangled name contains `template_args`.
    An instruction belongs to the kernel-body line it was inlined into (last entry of its inline chain) and to the region whose
    marker precedes that line; inside the triangle iteration, instructions inlined from triangle_test follow its own markers.
    ptxas unrolls the per-leaf triangle loop: the static count of those regions is divided by the number of copies (global
    loads found / the two 256-bit loads one iteration makes)."""
    kernel_marks, tri_marks, (tri_first, tri_last) = region_lines(names)
    tri_region = names.index("COST_TRI")
    tri_family = {tri_region, names.index("COST_TRI_HIT")} | {r for _, r in tri_marks}
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", OBJ], cwd=tmp, check=True, capture_output=True)
        cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
        text = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True, check=True).stdout
    sections = re.split(r"\n//-+ \.text\.", text)
    sec = [s for s in sections if s.startswith("_ZN3rdn") and "k_trace_ordered_rounds" in s.split("\n")[0] and template_args in s.split("\n")[0]]
    if len(sec) != 1:
        raise RuntimeError(f"{len(sec)} instantiations match {template_args}")
    counts = [0.0] * len(names)
    total = 0
    tri_loads = 0
    chain: list[int] = []   # traverse.cu lines of the current inline chain, innermost first
    chain_open = False
    for line in sec[0].split("\n"):
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            if not chain_open:
                chain, chain_open = [], True
            for f, l in ((m.group(1), m.group(2)), (m.group(3), m.group(4))):
                if f and f.endswith("traverse.cu"):
                    chain.append(int(l))
                elif f and f.endswith("ordered_rounds.inc"):
                    chain.append(INC_AT + int(l) * 1e-5)
            continue
        mi = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not mi:
            continue
        chain_open = False
        total += 1
        outer = chain[-1] if chain else kernel_marks[0][0]
        region = 0
        for mark_line, r in kernel_marks:
            if mark_line <= outer:
                region = r
        if region == tri_region:
            inner = next((l for l in chain if tri_first <= l <= tri_last), None)
            if inner is not None:
                for mark_line, r in tri_marks:
                    if mark_line <= inner:
                        region = r
            if mi.group(1).startswith("LDG"):
                tri_loads += 1
        counts[region] += 1
        DUMP.setdefault(region, []).append(f"{outer:10.5f} {line.strip()}")
    copies = max(1, round(tri_loads / 2))
    for r in tri_family:
        counts[r] /= copies
    return counts, total


CONFIGS = {
    # name: (scene builder, width, height, ray flags, tmin)
    "c1": ("sphere", 64, 1024, 1024, 0x10, 0.0),
    "c2": ("torus", 708, 1920, 1080, 0x10, 0.01),     # BASELINE config 2 (a few minutes on the emulator)
    "c2s": ("torus", 708, 960, 540, 0x10, 0.01),      # config 2's scene, a quarter of the pixels
    "c2t": ("torus", 256, 640, 360, 0x10, 0.01),
    "c3": ("torus_bounce", 708, 1920, 1080, 0x00, 0.01),  # BASELINE config 3: one cosine-weighted bounce ray per primary hit of config 2
    "c4": ("instances", 224, 1920, 1080, 0x10, 0.0),  # BASELINE config 4: 10,000 instances of a 100,352-triangle sphere
}


def dynamic_counts(cfg: str, names: list[str]):
    import numpy as np
    os.environ["RDN_SIMT_COST"] = "1"
    import build_emu
    from rendiation_b200 import api, scenes as S
    api.LIB_PATH = build_emu.build()
    import helpers
    kind, seg, w, h, flags, tmin = CONFIGS[cfg]
    if kind == "sphere":
        sp, _ = helpers.sphere_c1(seg=seg)
        rays = S.pinhole_rays(w, h, tmin, 100.0)
    elif kind == "instances":
        pos, idx = S.uv_sphere_mesh(seg, seg)
        sp = helpers.ScenePair()
        b = sp.blas([(pos, idx.reshape(-1), 1)])
        sp.bind([sp.tlas(S.instance_grid(100, 100, b, 3.5, -200.0))])
        sp.build()
        rays = S.pinhole_rays(w, h, tmin, 1000.0, aspect_correct=True)
    elif kind == "torus_bounce":
        sp, (pos, idx, m) = helpers.torus_scene(seg)
        primary = S.pinhole_rays(w, h, tmin, 100.0, aspect_correct=True)
        first, _ = sp.o.trace(primary, ray_flags=0x10, n_threads=os.cpu_count() or 4)
        d = np.stack([primary["dx"], primary["dy"], primary["dz"]], -1)
        hit = first["instance_id"] != 0xFFFFFFFF
        normals = np.zeros((primary.shape[0], 3), np.float32)
        normals[hit] = S.geometric_normals(pos, idx, first["primitive_id"][hit], m, d[hit])
        rays, _ = S.bounce_rays(primary, first, normals)
        w = 0  # a ray list, not a grid
    else:
        sp, _ = helpers.torus_scene(seg)
        rays = S.pinhole_rays(w, h, tmin, 100.0, aspect_correct=True)
    L = ctypes.CDLL(api.LIB_PATH)
    L.simt_cost_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    def aligned(n, dtype):  # device arrays must be 32-byte aligned (256-bit accesses)
        raw = np.zeros(n * 32 + 64, np.uint8)
        off = (-raw.ctypes.data) % 64
        return raw[off:off + n * 32].view(dtype)

    hits = aligned(rays.shape[0], api.HIT_DTYPE)
    src = rays
    rays = aligned(src.shape[0], api.RAY_DTYPE)
    rays[:] = src
    # device-resident call: one launch over the whole frame, as in bench.py's `value`
    sp.p.trace_closest_device(rays.ctypes.data, rays.shape[0], hits.ctypes.data, ray_flags=flags, grid_width=w)  # warm (scratch allocation)
    L.simt_cost_reset()
    t0 = time.time()
    sp.p.trace_closest_device(rays.ctypes.data, rays.shape[0], hits.ctypes.data, ray_flags=flags, grid_width=w)
    dt = time.time() - t0
    warp = (ctypes.c_uint64 * len(names))()
    lane = (ctypes.c_uint64 * len(names))()
    L.simt_cost_read(warp, lane, len(names))
    want, _ = sp.o.trace(rays, ray_flags=flags, n_threads=os.cpu_count() or 4)
    same = helpers.identical_hits(hits, want)
    return list(warp), list(lane), rays.shape[0], dt, bool(same)


def mangled(k, drain=True, ld256=True, wide4=False, inst_loop=False, share=False, anyhit=False, history=False, topup=False):
    """mangled template arguments <K, MINB, DRAIN_TIES, IRREGULAR, LD256, HOT, WIDE4, INST_LOOP, SHARE, ANYHIT, HISTORY, TOPUP> of an instantiation"""
    bl = lambda v: "Lb1E" if v else "Lb0E"
    return f"ILi{k}ELi8E{bl(drain)}Lb0E{bl(ld256)}Lb0E{bl(wide4)}{bl(inst_loop)}Li{int(share)}E{bl(anyhit)}{bl(history)}{bl(topup)}E"  # share: 0 never, 1 always, 2 late


def main():
    cfg = next((a for a in sys.argv[1:] if a in CONFIGS), "c2s")
    # the instantiation RDN_ORDERED_VARIANT selects (grids; ray lists take the sharing instantiation by default)
    variants = {0: mangled(3, inst_loop=True), 2: mangled(2), 9: mangled(3, drain=False, inst_loop=True), 30: mangled(3, ld256=False, inst_loop=True),
                60: mangled(2, wide4=True), 61: mangled(1, wide4=True), 100: mangled(3, inst_loop=True), 110: mangled(3, inst_loop=True, share=True)}
    variant = int(sys.argv[sys.argv.index("--variant") + 1]) if "--variant" in sys.argv else 0
    template_args = variants[variant]
    if variant:
        os.environ["RDN_ORDERED_VARIANT"] = str(variant)  # read by the library at its first launch (and by the flattener: wide4 view)
    names = region_names()
    static, total_static = static_counts(template_args, names)
    if "--dump" in sys.argv:
        want = "COST_" + sys.argv[sys.argv.index("--dump") + 1].upper()
        print("\n".join(DUMP.get(names.index(want), [])))
        return
    warp, lane, n_rays, dt, same = dynamic_counts(cfg, names)
    rows = []
    model_total = 0
    lane_total = 0
    for i, name in enumerate(names):
        issued = int(round(static[i] * warp[i]))
        model_total += issued
        lane_total += static[i] * lane[i]
        rows.append({"region": name[5:].lower(), "sass_instructions": round(static[i], 1), "warp_issues": warp[i], "lane_passes": lane[i],
                     "lanes_per_issue": round(lane[i] / warp[i], 2) if warp[i] else None, "warp_instructions": issued})
    for r in rows:
        r["share"] = round(r["warp_instructions"] / model_total, 4) if model_total else 0
    out = {"config": cfg, "variant": variant, "rays": n_rays, "emulation_seconds": round(dt, 1), "result_identical_to_oracle": same,
           "sass_instructions_in_kernel": total_static, "model_warp_instructions": model_total,
           "model_warp_instructions_per_ray": round(model_total / n_rays, 2),
           "model_active_lanes": round(lane_total / model_total, 2) if model_total else None, "regions": rows}
    if cfg == "c2":
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_r1f_sass_mix_k_trace_ordered_c2.json")))
        out["ncu_warp_instructions"] = ncu["warp_instructions"]
        out["ncu_active_lanes"] = round(ncu["avg_active_lanes"], 2)
        out["model_over_ncu"] = round(model_total / ncu["warp_instructions"], 3)
    if "--json" in sys.argv:
        print(json.dumps(out, indent=1))
        return
    print(json.dumps({k: v for k, v in out.items() if k != "regions"}))
    for r in rows:
        print(f"{r['region']:15s} sass/pass {r['sass_instructions']:7.1f}  warp issues {r['warp_issues']:9d}  lanes/issue "
              f"{(r['lanes_per_issue'] or 0):5.2f}  warp instr {r['warp_instructions']:10d}  {100 * r['share']:5.1f} %")


if __name__ == "__main__":
    main()
