"""Generate tests/golden/*.npz from the ORACLE (oracle/, the C restatement of the reference CPU path).

The reference itself cannot run here (no rustc/cargo) and holds no golden vectors for traversal, so
these fixtures pin the oracle's output ("parity unpinned" upstream): they freeze it against drift and give the
GPU tests a committed target that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import helpers  # noqa: E402
from rendiation_b200 import scenes as S  # noqa: E402


def main():
    # 1. the reference's own fixture scenes (geometry/naive/test.rs:9-225) under its 256x256-style pinhole, at 64x64
    sp, handles = helpers.reference_fixture(product=False)
    rays = S.pinhole_rays(64, 64, 0.0, 100.0)
    out = {"rays": rays}
    for k in range(5):
        for name, flags in (("cull_back", 0x10), ("none", 0x00), ("first_hit", 0x04 | 0x10)):
            hits, ctr = sp.o.trace(rays, ray_flags=flags, tlas_idx=k)
            out[f"hits_tlas{k}_{name}"] = hits
            out[f"ctr_tlas{k}_{name}"] = np.array([ctr[c] for c in ("bvh_visit", "bvh_hit", "tri_visit", "tri_hit", "inst_visit", "ref_abort")], np.uint64)
    np.savez_compressed(os.path.join(HERE, "reference_fixture_64.npz"), **out)

    # 2. BASELINE config 1 at reduced ray count: 64x64-segment sphere x5 at z=-10, 96x96 rays, path A and path B
    sp, (pos, idx, m) = helpers.sphere_c1(product=False)
    rays = S.pinhole_rays(96, 96, 0.0, 100.0)
    hits_b, ctr = sp.o.trace(rays, ray_flags=0x10)
    import oracle
    wpos = S.mat4_apply_point(m, pos)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([wpos[tri].min(1), wpos[tri].max(1)], axis=1)
    bvh = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, 50, 2)
    hits_a = bvh.query_nearest(wpos, idx, rays, oracle.FACE_DOUBLE)
    np.savez_compressed(os.path.join(HERE, "c1_sphere_96.npz"), rays=rays, hits_b=hits_b, hits_a=hits_a,
                        ctr=np.array([ctr[c] for c in ("bvh_visit", "bvh_hit", "tri_visit", "tri_hit", "inst_visit", "ref_abort")], np.uint64))
    # 3. the file the reference's own test writes: test_cpu_triangle (geometry/naive/test.rs:234-299) traces a 256x256 pinhole
    # grid over tlas 0 of the fixture with CULL_BACK_FACING and dumps `primitive_idx % 12 + 1` (0 = miss) as trace_cpu.pbm, and
    # prints the four visit counters.  A maintainer with the Rust toolchain can run that test and byte-compare the two files:
    # the one place where this oracle can be pinned against the real reference.
    sp, handles = helpers.reference_fixture(product=False)
    text, counters = trace_cpu_pbm(sp.o)
    open(os.path.join(HERE, "trace_cpu.pbm"), "w").write(text)
    import json
    json.dump(counters, open(os.path.join(HERE, "trace_cpu_counters.json"), "w"), indent=1)
    print("golden written")


def trace_cpu_pbm(scene, trace=None):
    """test_cpu_triangle's output: (the text of trace_cpu.pbm, its counter printout).  `trace(rays, ray_flags, tlas_idx)` defaults
    to the oracle scene's own traversal; the GPU tests pass the product's."""
    W = H = 256
    PRIMITIVE_IDX_MAX = 12
    rays = S.pinhole_rays(W, H, 0.0, 100.0)
    if trace is None:
        hits, ctr = scene.trace(rays, ray_flags=0x10, tlas_idx=0)
    else:
        hits, ctr = trace(rays, 0x10, 0)
    ids = np.where(hits["instance_id"] != 0xFFFFFFFF, hits["primitive_id"] % PRIMITIVE_IDX_MAX + 1, 0).reshape(H, W)
    text = f"P2\n{W} {H}\n{PRIMITIVE_IDX_MAX}\n" + "".join(" ".join(str(int(v)) for v in row) + "\n" for row in ids)
    counters = {"tri visit count": int(ctr["tri_visit"]), "tri hit count": int(ctr["tri_hit"]), "bvh visit count": int(ctr["bvh_visit"]),
                "bvh hit count": int(ctr["bvh_hit"])}
    return text, counters


if __name__ == "__main__":
    main()
