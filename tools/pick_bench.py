"""f2 measurement: brute-force picks per second on the device vs the oracle on the host cores (1 M-triangle torus).
Usage: python tools/pick_bench.py [n_rays]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from rendiation_b200 import api, scenes as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pos, idx = S.torus_mesh(708, 708, 1.0, 0.35)
idx = idx.reshape(-1)
rng = np.random.default_rng(1)
o = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
t = pos[rng.integers(0, pos.shape[0], n)]
d = (t - o).astype(np.float64); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
rays = S.make_rays(o, d, 0.0, 1e30)
mesh = api.PickMesh(pos, idx, api.TOPOLOGY_TRIANGLE_LIST)
mesh.ray_intersect_nearest(rays[:8])
out = {}
for m in (1, 16, n):
    t0 = time.perf_counter()
    reps = 20 if m < n else 5
    for _ in range(reps):
        got = mesh.ray_intersect_nearest(rays[:m])
    dt = (time.perf_counter() - t0) / reps
    out[f"gpu_batch{m}"] = {"ms_per_call": dt * 1e3, "picks_per_s": m / dt, "primitive_tests_per_s": m * mesh.primitive_count / dt}
cores = os.cpu_count() or 1
m = min(n, 64)
t0 = time.perf_counter(); want = oracle.pick_nearest(pos, idx, api.TOPOLOGY_TRIANGLE_LIST, rays[:m], n_threads=cores); dt = time.perf_counter() - t0
out["cpu_oracle"] = {"cores": cores, "rays": m, "picks_per_s": m / dt, "primitive_tests_per_s": m * mesh.primitive_count / dt}
t0 = time.perf_counter(); oracle.pick_nearest(pos, idx, api.TOPOLOGY_TRIANGLE_LIST, rays[:4], n_threads=1); dt = time.perf_counter() - t0
out["cpu_oracle_single_thread"] = {"picks_per_s": 4 / dt}
out["bit_identical"] = bool(mesh.ray_intersect_nearest(rays[:m]).tobytes() == want.tobytes())
out["workload"] = f"ray_intersect_nearest over a {mesh.primitive_count}-triangle torus (brute force, every primitive per pick), host ray / hit buffers"
print(json.dumps(out))
