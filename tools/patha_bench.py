"""BASELINE configs[0]: 1M coherent primary rays vs the 64x64-segment sphere (8,192 triangles) through the content/space BVH
(path A: FlattenBVH + intersect_nearest_bvh).  Times the oracle on the host cores and the device-resident query on cuda:0
and checks the results bit for bit.   python tools/patha_bench.py [iters]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from rendiation_b200 import api, scenes as S  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pos, idx = S.uv_sphere_mesh(64, 64)
wpos = S.mat4_apply_point(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), pos)
tri = idx.reshape(-1, 3)
boxes = np.concatenate([wpos[tri].min(1), wpos[tri].max(1)], 1)
rays = S.pinhole_rays(1024, 1024, 0.0, 100.0)
n = rays.shape[0]
cores = os.cpu_count() or 1
out = []
for name, opt in (("SAH(4) depth 50 bin 2", (50, 2)), ("SAH(4) default depth 10 bin 50", (10, 50))):
    ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, *opt)
    t0 = time.perf_counter(); want = ob.query_nearest(wpos, idx, rays, api.FACE_DOUBLE, cores); t_cpu = time.perf_counter() - t0
    sub = rays[::16].copy()
    t0 = time.perf_counter(); ob.query_nearest(wpos, idx, sub, api.FACE_DOUBLE, 1); t_cpu1 = time.perf_counter() - t0
    pb = api.build_bvh_for_abstract_mesh(wpos, idx, api.SAH(4), api.TreeBuildOption(*opt))
    api.upload_bvh(pb, wpos, idx, 0)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
    d_out = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        api.intersect_nearest_bvh_device(pb, d_rays.data_ptr(), n, d_out.data_ptr(), api.FACE_DOUBLE, st)
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); api.intersect_nearest_bvh_device(pb, d_rays.data_ptr(), n, d_out.data_ptr(), api.FACE_DOUBLE, st); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    got = d_out.cpu().numpy().view(api.MESH_HIT_DTYPE).reshape(-1)
    t0 = time.perf_counter(); got_host = api.intersect_nearest_bvh(wpos, idx, rays, pb, api.FACE_DOUBLE); t_host = time.perf_counter() - t0
    out.append({"bvh": name, "rays": n, "triangles": int(tri.shape[0]), "gpu_resident_mrays": n / np.mean(ms) / 1e3, "gpu_resident_ms": float(np.mean(ms)),
                "gpu_host_buffers_mrays": n / t_host / 1e6, "cpu_oracle_mrays": n / t_cpu / 1e6, "cpu_cores": cores,
                "cpu_single_thread_mrays": sub.shape[0] / t_cpu1 / 1e6, "bit_identical": bool(got.tobytes() == want.tobytes() == got_host.tobytes()),
                "hits": int(want["hit"].sum())})
for o in out:
    print(json.dumps(o))
