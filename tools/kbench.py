"""Kernel experiment harness (not part of the product): times the device-resident trace on a BASELINE config and
checks a row-strided sample against the oracle.  Usage: [RDN_ORDERED_VARIANT=k] python tools/kbench.py [c1|c2|c3|c4] [iters]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from rendiation_b200 import api, scenes as S  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
check = os.environ.get("KBENCH_CHECK", "1") == "1"

T, Sc, Rx, mul = S.mat4_translate, S.mat4_scale, S.mat4_rotate_x, S.mat4_mul
sysm = api.NaiveSahBVHSystem()
osc = oracle.Scene()
flags = 0x10
if cfg in ("c2", "c3", "c2i"):
    pos, idx = S.torus_mesh(708, 708, 1.0, 0.35)
    m = mul(mul(T(0, 0, -10), Sc(5, 5, 5)), Rx(-0.5))
    inst_o = inst_p = None
    W, H = 1920, 1080
    rays = S.pinhole_rays(W, H, 0.01, 100.0, aspect_correct=True)
    insts = lambda b: S.make_instance(m, b)
elif cfg == "c1":
    pos, idx = S.uv_sphere_mesh(64, 64)
    m = mul(T(0, 0, -10), Sc(5, 5, 5))
    W, H = 1024, 1024
    rays = S.pinhole_rays(W, H, 0.0, 100.0)
    insts = lambda b: S.make_instance(m, b)
elif cfg == "c4":
    pos, idx = S.uv_sphere_mesh(224, 224)
    W, H = 1920, 1080
    rays = S.pinhole_rays(W, H, 0.0, 1000.0, aspect_correct=True)
    insts = lambda b: S.instance_grid(100, 100, b, 3.5, -200.0)
extra = None
if cfg == "c2i":  # config 2 plus a few cubes, one of them under a singular transform (an irregular instance: rays whose range meets
    # its box are handed to the reference-order walk); it covers a corner of the frame
    def extra(cube):
        return np.concatenate([S.make_instance(mul(T(-7.0, 3.5, -12.0), Sc(2.0, 0.0, 2.0)), cube),
                               S.make_instance(mul(T(7.0, 3.5, -12.0), Sc(1.5, 1.5, 1.5)), cube),
                               S.make_instance(mul(T(7.0, -3.5, -12.0), Sc(1.5, 1.5, 1.5)), cube)])
t0 = time.time()
b = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
if extra is not None:
    cube = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(S.CUBE_POSITION, S.CUBE_INDEX)])
    _insts = insts
    insts = lambda bb: np.concatenate([_insts(bb), extra(1)])
t = sysm.create_top_level_acceleration_structure(insts(b.id))
sysm.bind_tlas([t]); sysm.commit()
t_build = time.time() - t0
ob = osc.create_blas([(pos, idx.reshape(-1), 1)])
if extra is not None:
    osc.create_blas([(S.CUBE_POSITION, S.CUBE_INDEX, 1)])
osc.bind_tlas([osc.create_tlas(insts(ob))]); assert osc.build() == 0
nt = os.cpu_count() or 4
grid = W
if cfg == "c3":  # incoherent cosine-weighted bounce rays off the primary hits, no culling
    ph = osc.trace(rays, ray_flags=0x10, n_threads=nt, want_counters=False)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], -1)
    hit = ph["instance_id"] != 0xFFFFFFFF
    normals = np.zeros((rays.shape[0], 3), np.float32)
    normals[hit] = S.geometric_normals(pos, idx, ph["primitive_id"][hit], m, d[hit])
    rays, _ = S.bounce_rays(rays, ph, normals)
    flags, grid = 0, 0
if os.environ.get("KBENCH_FLAGS"):  # e.g. 0x14: ACCEPT_FIRST_HIT_AND_END_SEARCH | CULL_BACK (shadow / AO rays)
    flags = int(os.environ["KBENCH_FLAGS"], 0)
rows = os.environ.get("KBENCH_ROWS")
if rows:
    r0, r1 = (int(x) for x in rows.split(":"))
    rays = rays.reshape(H, W)[r0:r1].reshape(-1).copy()
rep = int(os.environ.get("KBENCH_REPEAT", "1"))
if rep > 1:
    rays = np.tile(rays, rep)
    grid = grid  # rows simply repeat: still a multiple of the width
n = rays.shape[0]
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
d_hits = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
if os.environ.get("KBENCH_SIDE_STREAM", "0") == "1":  # a created stream instead of the legacy default stream
    _side = torch.cuda.Stream()
    torch.cuda.set_stream(_side)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    sysm.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=flags, grid_width=grid, stream=st)
stats = sysm.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=flags, grid_width=grid, stream=st, want_stats=True)
ms = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sysm.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=flags, grid_width=grid, stream=st)
    e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = np.array(ms)
# back-to-back launches on one stream (no flush, no events in between): consecutive launches may overlap their tails.
# Every launch writes its OWN hit buffer (pre-filled with 0xAB) so that a ray skipped by an overlapped launch cannot hide
# behind the result of another launch; all buffers are compared with the serialised result below.
b2b_hits = [torch.full((n, 32), 0xAB, dtype=torch.uint8, device="cuda") for _ in range(iters)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for k in range(iters):
    sysm.trace_closest_device(d_rays.data_ptr(), n, b2b_hits[k].data_ptr(), ray_flags=flags, grid_width=grid, stream=st, overlap_previous=True)
e1.record(); torch.cuda.synchronize()
b2b_ms = e0.elapsed_time(e1) / iters
b2b_ok = all(torch.equal(h, d_hits) for h in b2b_hits)
del b2b_hits
# the same with an event recorded between launches (does a marker between two kernels serialise them?)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
evs[0].record()
for k in range(iters):
    sysm.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=flags, grid_width=grid, stream=st)
    evs[k + 1].record()
torch.cuda.synchronize()
b2b_ev_ms = evs[0].elapsed_time(evs[-1]) / iters
ok = None
if check:
    sel = np.arange(0, n, 7)
    want = osc.trace(rays[sel], ray_flags=flags, n_threads=nt, want_counters=False)
    got = d_hits.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)[sel]
    ok = got.tobytes() == want.tobytes()
print(f"cfg={cfg} flags={flags:#x} variant={os.environ.get('RDN_ORDERED_VARIANT','default')} skip_tie={'RDN_DEBUG_SKIP_TIE' in os.environ} rays={n} "
      f"irregular={sysm.build_stats()['irregular_instances']}/{sysm.build_stats()['reference_routed_tlas']} mean_ms={ms.mean():.4f} min_ms={ms.min():.4f} Mrays/s(mean)={n/ms.mean()/1e3:.1f} best={n/ms.min()/1e3:.1f} ties={stats['tie_rays']} "
      f"bit_identical_sample={ok} build_s={t_build:.2f} | back-to-back {n/b2b_ms/1e3:.1f} Mrays/s ({b2b_ms:.4f} ms, all {iters} results identical to the serialised one: {b2b_ok}), with events between {n/b2b_ev_ms/1e3:.1f} "
      f"pdl={os.environ.get('RDN_PDL','1')} side_stream={os.environ.get('KBENCH_SIDE_STREAM','0')}")
