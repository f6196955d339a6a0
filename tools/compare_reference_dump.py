"""Compare what the REAL reference produced (tools/pin_against_reference.sh) with this repository's oracle.

    python tools/compare_reference_dump.py --dump b200_fixture_dump.bin [--pbm trace_cpu.pbm] [--log test_cpu_triangle.log]

--dump : rust/pin-against-reference/test_dump_b200.rs output.  The fixture is rebuilt in oracle/ FROM THE DUMPED INPUTS (positions,
         indices, instance records), the dumped ray directions are traced, and every record of the 5 TLASes x 3 flag sets must
         agree: hit / miss, geometry_idx, primitive_idx exactly, the distance bit for bit; the four visit counters too.
--pbm  : the reference's own trace_cpu.pbm against tests/golden/trace_cpu.pbm (byte comparison).  The golden was generated from
         numpy's tessellation of the fixture; if Rust's sin / cos round differently on a vertex the files can differ on a
         silhouette pixel while --dump (same inputs on both sides) still agrees — the tool says which.
--log  : the counters test_cpu_triangle printed, against tests/golden/trace_cpu_counters.json.
Exit status 0 = pinned.  `--self-test` writes a dump from the oracle itself in the reference's format and reads it back (CPU test)."""
from __future__ import annotations

import argparse
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
FLAG_SETS = (("cull_back", 0x10), ("none", 0x00), ("first_hit", 0x04 | 0x10))
FAR = 100.0


def read_sections(path):
    data = open(path, "rb").read()
    assert data[:8] == b"RDNDUMP1", "not a b200 fixture dump"
    off, out = 8, []
    sizes = {"positions": 12, "indices": 4, "geom_flags": 4, "aabbs": 24, "ray_dirs": 12}
    while off < len(data):
        name = data[off:off + 16].rstrip(b"\0").decode(); off += 16
        (count,) = struct.unpack_from("<Q", data, off); off += 8
        if name == "blas":
            nbytes = 4
        elif name == "tlas":
            nbytes = 4 + count * (64 + 20)
        elif name[0] == "t" and "_" in name:
            nbytes = count * 16
        elif name[0] == "c" and "_" in name:
            nbytes = count * 4
        else:
            nbytes = count * sizes[name]
        out.append((name, count, data[off:off + nbytes])); off += nbytes
    return out


def scene_from_dump(sections):
    import oracle
    from rendiation_b200 import scenes as S
    osc = oracle.Scene()
    blas, cur = [], None
    tlas = []
    results, counters, dirs = {}, {}, None
    pending_pos = None
    for name, count, raw in sections:
        if name == "blas":
            cur = []
            blas.append(cur)
        elif name == "positions":
            pending_pos = np.frombuffer(raw, "<f4").reshape(-1, 3).copy()
        elif name == "indices":
            cur.append([pending_pos, np.frombuffer(raw, "<u4").copy(), 0])
        elif name == "aabbs":
            cur.append([np.frombuffer(raw, "<f4").reshape(-1, 6).copy(), None, 0, True])
        elif name == "geom_flags":
            cur[-1][2] = int(np.frombuffer(raw, "<u4")[0])
        elif name == "tlas":
            rec = np.frombuffer(raw[4:], np.dtype([("m", "<f4", 16), ("u", "<u4", 5)]))
            inst = np.zeros(count, S.INSTANCE_DTYPE)
            inst["transform"] = rec["m"]
            inst["instance_custom_index"], inst["mask"] = rec["u"][:, 0], rec["u"][:, 1]
            inst["sbt_offset"], inst["flags"], inst["blas_handle"] = rec["u"][:, 2], rec["u"][:, 3], rec["u"][:, 4]
            tlas.append(inst)
        elif name == "ray_dirs":
            dirs = np.frombuffer(raw, "<f4").reshape(-1, 3).copy()
        elif name[0] == "t":
            results[name[1:]] = np.frombuffer(raw, "<u4").reshape(-1, 4).copy()
        elif name[0] == "c":
            counters[name[1:]] = np.frombuffer(raw, "<u4").copy()
    for geoms in blas:
        osc.create_blas([tuple(g) for g in geoms])
    handles = [osc.create_tlas(t) for t in tlas]
    osc.bind_tlas(handles)
    assert osc.build() == 0
    rays = np.zeros(dirs.shape[0], S.RAY_DTYPE)
    rays["tmin"], rays["tmax"] = 0.0, FAR
    rays["dx"], rays["dy"], rays["dz"] = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    return osc, rays, results, counters, len(tlas)


def compare_dump(path) -> bool:
    osc, rays, results, counters, n_tlas = scene_from_dump(read_sections(path))
    ok = True
    for k in range(n_tlas):
        for name, flags in FLAG_SETS:
            key = f"{k}_{name}"
            hits, ctr = osc.trace(rays, ray_flags=flags, tlas_idx=k)
            ref = results[key]
            hit = hits["instance_id"] != 0xFFFFFFFF
            same = (np.array_equal(ref[:, 0] != 0, hit) and np.array_equal(ref[hit, 1], hits["geometry_id"][hit]) and
                    np.array_equal(ref[hit, 2], hits["primitive_id"][hit]) and np.array_equal(ref[hit, 3], hits["t"][hit].view(np.uint32)))
            c = counters[key]
            same_c = [int(c[0]), int(c[1]), int(c[2]), int(c[3])] == [ctr["tri_visit"], ctr["tri_hit"], ctr["bvh_visit"], ctr["bvh_hit"]]
            print(f"tlas {k} flags {name:9s}: {int(hit.sum()):5d} hits  records {'identical' if same else 'DIFFER'}  visit counters {'identical' if same_c else 'DIFFER'}")
            ok = ok and same and same_c
    return ok


def write_dump_from_oracle(path):
    """the reference test's file, produced by the oracle on the repository's own rendering of the fixture (self-test of the format)"""
    import helpers
    from rendiation_b200 import scenes as S
    sp, handles = helpers.reference_fixture(product=False)
    out = bytearray(b"RDNDUMP1")

    def section(name, count, payload):
        out.extend(name.encode().ljust(16, b"\0")); out.extend(struct.pack("<Q", count)); out.extend(payload)

    for b, geoms in enumerate(sp.blas_sources):
        section("blas", len(geoms), struct.pack("<I", b))
        for g in geoms:
            section("positions", g[0].shape[0], np.ascontiguousarray(g[0], "<f4").tobytes())
            section("indices", g[1].size, np.ascontiguousarray(g[1], "<u4").tobytes())
            section("geom_flags", 1, struct.pack("<I", g[2]))
    for t, inst in enumerate(sp.tlas_sources):
        payload = bytearray(struct.pack("<I", t))
        for i in inst:
            payload.extend(np.ascontiguousarray(i["transform"], "<f4").tobytes())
            payload.extend(struct.pack("<5I", int(i["instance_custom_index"]), int(i["mask"]), int(i["sbt_offset"]),
                                       int(i["flags"]), int(i["blas_handle"])))
        section("tlas", len(inst), payload)
    rays = S.pinhole_rays(64, 64, 0.0, FAR)
    dirs = np.stack([rays["dx"], rays["dy"], rays["dz"]], -1).astype("<f4")
    section("ray_dirs", dirs.shape[0], dirs.tobytes())
    for k in range(len(sp.tlas_sources)):
        for name, flags in FLAG_SETS:
            hits, ctr = sp.o.trace(rays, ray_flags=flags, tlas_idx=k)
            hit = hits["instance_id"] != 0xFFFFFFFF
            rec = np.zeros((rays.shape[0], 4), "<u4")
            rec[:, 0] = hit
            rec[:, 1] = np.where(hit, hits["geometry_id"], 0xFFFFFFFF)
            rec[:, 2] = np.where(hit, hits["primitive_id"], 0xFFFFFFFF)
            rec[:, 3] = np.where(hit, hits["t"].view(np.uint32), np.float32(FAR).view(np.uint32))
            section(f"t{k}_{name}", rays.shape[0], rec.tobytes())
            section(f"c{k}_{name}", 4, struct.pack("<4I", ctr["tri_visit"], ctr["tri_hit"], ctr["bvh_visit"], ctr["bvh_hit"]))
    open(path, "wb").write(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dump"); ap.add_argument("--pbm"); ap.add_argument("--log"); ap.add_argument("--self-test", action="store_true")
    a = ap.parse_args()
    ok = True
    if a.self_test:
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            p = os.path.join(tmp, "b200_fixture_dump.bin")
            write_dump_from_oracle(p)
            ok = compare_dump(p)
    if a.dump:
        ok = compare_dump(a.dump) and ok
    if a.pbm:
        same = open(a.pbm, "rb").read() == open(os.path.join(ROOT, "tests", "golden", "trace_cpu.pbm"), "rb").read()
        print("trace_cpu.pbm:", "identical to tests/golden/trace_cpu.pbm" if same else
              "DIFFERS from tests/golden/trace_cpu.pbm (if --dump agrees, the difference is the mesh generator's sin / cos, not the traversal)")
        ok = ok and (same or bool(a.dump))
    if a.log:
        want = json.load(open(os.path.join(ROOT, "tests", "golden", "trace_cpu_counters.json")))
        text = open(a.log).read()
        got = {k: int(text.split(k + ":")[1].split()[0]) for k in want if k + ":" in text}
        print("visit counters:", "identical" if got == want else f"DIFFER {got} vs {want}")
        ok = ok and (got == want or bool(a.dump))
    print("PINNED: the oracle reproduces the reference on its own fixture" if ok else "NOT PINNED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
