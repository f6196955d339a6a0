"""Shader-binding-table dispatch after a trace (first slice) against the numpy restatement of
trace_task.rs:206-268 / api/ctx.rs:53-55 / sbt.rs (oracle/sbt.py).  Integer work: bit-exact.

(The file name sorts last on purpose: these kernels are the newest and `pytest -x` should reach everything else first.)"""
import numpy as np
import pytest

import helpers
from oracle import sbt as osbt
from rendiation_b200 import api, scenes as S

pytestmark = pytest.mark.gpu


def _scene():
    """Two BLASes, five instances with different SBT record offsets.  The multi-geometry BLAS is created last and its first
    geometry is a backdrop quad that spans the others: the reference takes an instance's box from the geometry whose global
    index equals the BLAS handle (naive/mod.rs blas_box), so only then are all three geometries reachable."""
    sp = helpers.ScenePair()
    spos, sidx = S.uv_sphere_mesh(24, 16)
    tpos, tidx = S.torus_mesh(24, 12, 1.0, 0.3)
    single = sp.blas([(tpos, tidx.reshape(-1), 1)])
    quad = np.float32([[-1.6, -1.6, -1.5], [1.6, -1.6, -1.5], [1.6, 1.6, -1.5], [-1.6, 1.6, -1.5]])
    multi = sp.blas([(quad, np.uint32([0, 1, 2, 0, 2, 3]), 1), (tpos, tidx.reshape(-1), 1), (spos * np.float32(0.45), sidx.reshape(-1), 1)])
    T, mul, Rx = S.mat4_translate, S.mat4_mul, S.mat4_rotate_x
    inst = np.concatenate([
        S.make_instance(mul(T(0, 0, -10), Rx(1.1)), multi, custom_index=7, sbt_offset=0),
        S.make_instance(mul(T(0, 3.5, -10), Rx(1.0)), single, custom_index=8, sbt_offset=6),
        S.make_instance(mul(T(0, -3.5, -10), Rx(1.3)), single, custom_index=9, sbt_offset=8),
        S.make_instance(mul(T(4.0, 2.5, -12), Rx(0.9)), multi, custom_index=10, sbt_offset=10),
        S.make_instance(mul(T(-4.0, -2.5, -9), Rx(1.2)), single, custom_index=11, sbt_offset=40),  # beyond the table: selects nothing
    ])
    sp.bind([sp.tlas(inst)])
    return sp.build()


def _tables(sp, ray_type_count=2):
    o = osbt.ShaderBindingTable(8, 2, ray_type_count)
    p = sp.p.create_sbt(8, 2, ray_type_count)
    o.config_ray_generation(3); p.config_ray_generation(3)
    rng = np.random.default_rng(5)
    for geometry_idx in range(8):
        for tlas_offset in range(2):
            for ray_ty in range(ray_type_count):
                closest = None if rng.random() < 0.25 else int(rng.integers(0, 3))
                if ray_ty + geometry_idx * ray_type_count + tlas_offset >= o.ray_hit.shape[0]:
                    continue
                o.config_hit_group(geometry_idx, tlas_offset, ray_ty, closest, 1, None)
                p.config_hit_group(geometry_idx, tlas_offset, ray_ty, api.HitGroupShaderRecord(closest, 1, None))
    o.config_missing(0, 1); p.config_missing(0, 1)
    return o, p


@pytest.mark.parametrize("cfg", [dict(), dict(sbt_ray_offset=1, sbt_ray_stride=2), dict(miss_index=1),
                                 dict(ray_flags=api.RAY_FLAG_SKIP_CLOSEST_HIT_SHADER), dict(sbt_ray_offset=3, sbt_ray_stride=5, miss_index=7)])
def test_dispatch_and_task_lists_match_the_restatement(cfg):
    sp = _scene()
    rays = S.pinhole_rays(192, 160, 0.01, 100.0)
    hits = sp.p.trace_closest_batch(rays, ray_flags=0, grid_width=192)
    want_hits, _ = sp.o.trace(rays, ray_flags=0, n_threads=4)
    assert hits.tobytes() == want_hits.tobytes()
    o, p = _tables(sp)
    assert p.ray_generation == 3
    sbt_offset = sp.p.arrays()["instances"]["sbt_offset"]
    want_task = osbt.dispatch(o, hits, sbt_offset, **cfg)
    want_queue, want_offsets = osbt.group(want_task, 3, 2)
    task, queue, offsets = p.dispatch(hits, 3, 2, **cfg)
    assert np.array_equal(task, want_task)
    assert np.array_equal(offsets, want_offsets) and np.array_equal(queue, want_queue)
    if not cfg:
        kinds = set(np.unique(task).tolist())
        assert {0, 1, 2, api.TASK_NONE, 1 | api.TASK_MISS_BIT} <= kinds, kinds   # every outcome occurs in the plain configuration
        assert offsets[-1] < rays.shape[0]                                        # ... so some rays spawn nothing


def test_device_resident_wave_and_edge_cases():
    import torch
    sp = _scene()
    o, p = _tables(sp)
    rays = S.pinhole_rays(96, 64, 0.01, 100.0)
    n = rays.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
    d_hits = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    d_task = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_queue = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(6, dtype=torch.int64, device="cuda")
    sp.p.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=0, grid_width=96, stream=st)
    p.dispatch_device(d_hits.data_ptr(), n, d_task.data_ptr(), stream=st)
    p.group_device(d_task.data_ptr(), n, 3, 2, d_queue.data_ptr(), d_off.data_ptr(), stream=st)
    torch.cuda.synchronize()
    hits = d_hits.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
    want_task = osbt.dispatch(o, hits, sp.p.arrays()["instances"]["sbt_offset"])
    want_queue, want_offsets = osbt.group(want_task, 3, 2)
    assert np.array_equal(d_task.cpu().numpy().view(np.uint32), want_task)
    off = d_off.cpu().numpy().view(np.uint64)
    assert np.array_equal(off, want_offsets)
    assert np.array_equal(d_queue.cpu().numpy().view(np.uint32)[:int(off[-1])], want_queue)
    # reconfiguring the table is picked up by the next dispatch
    p.config_missing(0, 0); o.config_missing(0, 0)
    p.dispatch_device(d_hits.data_ptr(), n, d_task.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(d_task.cpu().numpy().view(np.uint32), osbt.dispatch(o, hits, sp.p.arrays()["instances"]["sbt_offset"]))
    # empty batch, no shaders asked for, records outside the table
    task, queue, offsets = p.dispatch(np.zeros(0, api.HIT_DTYPE), 3, 2)
    assert task.size == 0 and queue.size == 0 and offsets.tolist() == [0] * 6
    task, queue, offsets = p.dispatch(hits, 0, 0)
    assert offsets.tolist() == [0] and queue.size == 0 and np.array_equal(task, osbt.dispatch(o, hits, sp.p.arrays()["instances"]["sbt_offset"]))
    p.config_hit_group(15, 0, 1, api.HitGroupShaderRecord(0))           # record 1 + 15*2 + 0 = 31: the last one of 8*2*2
    with pytest.raises(api.RdnError):
        p.config_hit_group(15, 1, 1, api.HitGroupShaderRecord(0))       # record 32: the reference's set_value(..).unwrap() panics
    with pytest.raises(api.RdnError):
        p.config_missing(2, 0)
