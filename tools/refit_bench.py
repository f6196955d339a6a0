"""TLAS-only update of BASELINE config 4 (10,000 instances of a 100,352-triangle sphere): rdn_rt_tlas_update + commit, against the
full commit of the same scene.  Prints one JSON line.   python tools/refit_bench.py [reps]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rendiation_b200 import api, scenes as S  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
pos, idx = S.uv_sphere_mesh(224, 224)
sysm = api.NaiveSahBVHSystem()
t0 = time.perf_counter()
b = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
t = sysm.create_top_level_acceleration_structure(S.instance_grid(100, 100, b.id, 3.5, -200.0))
sysm.bind_tlas([t]); sysm.commit()
full_ms = (time.perf_counter() - t0) * 1e3
rays = S.pinhole_rays(1920, 1080, 0.0, 1000.0, aspect_correct=True)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
d_hits = torch.zeros_like(d_rays)
sysm.trace_closest_device(d_rays.data_ptr(), rays.shape[0], d_hits.data_ptr(), ray_flags=0x10, grid_width=1920)
torch.cuda.synchronize()
ms = []
for k in range(reps):
    inst = S.instance_grid(100, 100, b.id, 3.5 + 0.01 * (k + 1), -200.0 - k)   # every instance moves
    t0 = time.perf_counter()
    sysm.update_top_level_acceleration_structure(t, inst)
    sysm.commit()
    ms.append((time.perf_counter() - t0) * 1e3)
    sysm.trace_closest_device(d_rays.data_ptr(), rays.shape[0], d_hits.data_ptr(), ray_flags=0x10, grid_width=1920)
    torch.cuda.synchronize()
st = sysm.build_stats()
print(json.dumps({"scene": "BASELINE configs[3]: 10,000 instances x 100,352 triangles", "full_commit_ms": full_ms, "tlas_update_commit_ms_median": float(np.median(ms)),
                  "tlas_update_commit_ms_min": float(min(ms)), "reps": reps, "tlas_only_commits": st["tlas_only_commits"], "last_bvh_build_ms": st["bvh_build_ms"],
                  "last_flatten_ms": st["flatten_ms"], "last_patch_ms": st["upload_ms"], "host_threads": os.cpu_count()}))
