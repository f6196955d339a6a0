"""PCIe probe (experiment harness): pinned H2D / D2H / both at once, then the host-buffer trace path (rdn_rt_trace_closest) on
BASELINE configs[1] at several RDN_HOST_CHUNK_RAYS.   python tools/pcie_probe.py [trace]"""
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def bandwidth():
    n = 66355200
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", lambda: (h2d(), d2h()))):
        t = timeit(fn)
        print(f"{name}: {t*1e3:.3f} ms for {n/1e6:.1f} MB each way -> {n/t/1e9:.1f} GB/s per direction", flush=True)


def trace_one():
    from rendiation_b200 import api, scenes as S
    pos, idx = S.torus_mesh(708, 708, 1.0, 0.35)
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
    rays = S.pinhole_rays(1920, 1080, 0.01, 100.0, aspect_correct=True)
    sysm = api.NaiveSahBVHSystem()
    b = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
    sysm.bind_tlas([sysm.create_top_level_acceleration_structure(S.make_instance(m, b.id))]); sysm.commit()
    n = rays.shape[0]
    h_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).pin_memory()
    h_hits = torch.zeros((n, 32), dtype=torch.uint8).pin_memory()
    t = timeit(lambda: sysm.trace_closest_host_ptr(h_rays.data_ptr(), n, h_hits.data_ptr(), ray_flags=0x10, grid_width=1920), reps=20)
    print(f"{t*1e3:.3f} ms/frame -> {n/t/1e6:.0f} Mrays/s e2e")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "trace_one":
        trace_one()
    else:
        bandwidth()
        if len(sys.argv) > 1 and sys.argv[1] == "trace":
            for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
                out = subprocess.run([sys.executable, __file__, "trace_one"], env=dict(os.environ, RDN_HOST_CHUNK_RAYS=str(chunk)),
                                     capture_output=True, text=True)
                print(f"chunk={chunk}: {out.stdout.strip()} {out.stderr.strip()[-300:]}", flush=True)
