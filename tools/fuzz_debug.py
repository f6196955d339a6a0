"""Debug harness for tests/test_gpu_fuzz.py: per profile / seed / trial, which kernel disagrees with the oracle, on what kind of ray.
Usage: python tools/fuzz_debug.py [profile ...] (default: all three), seeds 0..7; FUZZ_SEED_BASE / FUZZ_SEEDS select others.
FUZZ_EMU=1 runs the campaign on the CPU emulation of the kernels (tests/simt) instead of a GPU."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from rendiation_b200 import api
if os.environ.get("FUZZ_EMU") == "1":
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
    import build_emu
    api.LIB_PATH = build_emu.build()
import helpers, test_gpu_fuzz as F

profiles = sys.argv[1:] or ["regular", "mixed", "hostile"]
total_bad = 0
for profile in profiles:
    for seed in range(int(os.environ.get("FUZZ_SEED_BASE", "0")), int(os.environ.get("FUZZ_SEED_BASE", "0")) + int(os.environ.get("FUZZ_SEEDS", "8"))):
        sp, n_tlas, rng = F._scene(1000 + seed, profile=profile)
        rays = F._rays(rng, 12000)
        st = sp.p.build_stats()
        arrays = sp.p.arrays()
        flag_sets = [0, 0x10, 0x20, 0x01, 0x40, 0x80 | 0x10, 0x100, 0x04, 0x02 | 0x20]
        for trial in range(5):
            flags = int(flag_sets[int(rng.integers(0, len(flag_sets)))])
            mask = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 0x3, 0xF0]))
            tlas_idx = int(rng.integers(0, n_tlas + 1))
            want, wctr = sp.o.trace(rays, ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx, n_threads=4)
            route_all, suspect = helpers.suspect_rays(arrays, rays, tlas_idx, mask)
            try:
                got = sp.p.trace_closest_batch(rays, ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx)
                err = None
            except api.RdnError as e:
                got, err = None, str(e)
            got_ref, gctr = sp.p.trace_counted(rays, ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx)
            ref_ok = helpers.identical_hits(got_ref, want)
            line = (f"{profile} seed={seed} trial={trial} flags={flags:#x} mask={mask:#x} tlas={tlas_idx}/{n_tlas} hits={(want['instance_id']!=0xFFFFFFFF).sum()} "
                    f"irregular tri/inst/routed={st['irregular_triangles']}/{st['irregular_instances']}/{st['reference_routed_tlas']} route_all={route_all} suspect={int(suspect.sum())} "
                    f"ref_abort={wctr['ref_abort']} reforder_ok={ref_ok} ctr_ok={gctr==wctr}")
            if err:
                print(line, "ORDERED ERROR:", err); total_bad += 1; continue
            g, w = helpers.canonical_nan(got), helpers.canonical_nan(want)
            bad = np.unique(np.nonzero(g.view(np.uint8).reshape(-1, 32) != w.view(np.uint8).reshape(-1, 32))[0])
            total_bad += bad.size + (0 if ref_ok else 1)
            print(line, f"ordered_mismatch={bad.size}")
            for i in bad[:3]:
                r = rays[i]
                print("   ray", i, "suspect" if suspect[i] else "kept", [float(r[k]) for k in r.dtype.names])
                print("   got ", [got[i][k].item() for k in got.dtype.names])
                print("   want", [want[i][k].item() for k in want.dtype.names])
print("TOTAL_BAD", total_bad)
