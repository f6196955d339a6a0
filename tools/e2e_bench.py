"""Host-buffer path (rdn_rt_trace_closest: rays in, hits out, chunked over 4 streams) on BASELINE configs[1], with page-locked caller
buffers and with ordinary (pageable) ones, which go through the library's staging pipeline.
Usage: [RDN_HOST_CHUNK_RAYS=n] [RDN_STAGE_THREADS=t] python tools/e2e_bench.py [iters]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rendiation_b200 import api, scenes as S

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pos, idx = S.torus_mesh(708, 708, 1.0, 0.35)
m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
sysm = api.NaiveSahBVHSystem()
b = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
sysm.bind_tlas([sysm.create_top_level_acceleration_structure(S.make_instance(m, b.id))]); sysm.commit()
W, H = 1920, 1080
rays = S.pinhole_rays(W, H, 0.01, 100.0, aspect_correct=True)
n = rays.shape[0]
h_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).pin_memory()
h_hits = torch.zeros((n, 32), dtype=torch.uint8).pin_memory()
d_rays = h_rays.cuda(); d_hits = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
sysm.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=0x10, grid_width=W, stream=torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    sysm.trace_closest_host_ptr(h_rays.data_ptr(), n, h_hits.data_ptr(), ray_flags=0x10, grid_width=W)
ts = []
for _ in range(iters):
    t0 = time.perf_counter()
    sysm.trace_closest_host_ptr(h_rays.data_ptr(), n, h_hits.data_ptr(), ray_flags=0x10, grid_width=W)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts)
same = bool(torch.equal(h_hits.cuda(), d_hits))
print(f"pinned   chunk={os.environ.get('RDN_HOST_CHUNK_RAYS','default')} mean_ms={ts.mean()*1e3:.4f} min_ms={ts.min()*1e3:.4f} "
      f"Mrays/s(mean)={n/ts.mean()/1e6:.1f} best={n/ts.min()/1e6:.1f} same_as_device_path={same}")
# ordinary memory, as a caller of the reference's API would hand over
p_rays = rays.view(np.uint8).reshape(-1, 32).copy()
p_hits = np.zeros((n, 32), np.uint8)
for _ in range(3):
    sysm.trace_closest_host_ptr(p_rays.ctypes.data, n, p_hits.ctypes.data, ray_flags=0x10, grid_width=W)
ts = []
for _ in range(iters):
    t0 = time.perf_counter()
    sysm.trace_closest_host_ptr(p_rays.ctypes.data, n, p_hits.ctypes.data, ray_flags=0x10, grid_width=W)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts)
same = bool(np.array_equal(p_hits, d_hits.cpu().numpy()))
print(f"pageable chunk={os.environ.get('RDN_HOST_CHUNK_RAYS','default')} threads={os.environ.get('RDN_STAGE_THREADS','default')} mean_ms={ts.mean()*1e3:.4f} "
      f"min_ms={ts.min()*1e3:.4f} Mrays/s(mean)={n/ts.mean()/1e6:.1f} best={n/ts.min()/1e6:.1f} same_as_device_path={same}")
