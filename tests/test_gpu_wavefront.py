"""rdn_rt_trace_ray — the wavefront executor in one call against oracle/wavefront.py, and the AO frame of
feature/ao.rs through that single entry against the stage-by-stage pipeline of test_gpu_raygen.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle
from oracle import sbt as OSBT
from oracle import wavefront as OW
from rendiation_b200 import api, scenes as S

sys.path.insert(0, os.path.dirname(__file__))
import helpers  # noqa: E402

pytestmark = pytest.mark.gpu
f32 = np.float32
EMU = os.environ.get("RDN_SIMT_EMU") == "1"


def dev_read(ptr, nbytes, stream=0):
    """device memory -> bytes, ordered behind what the stream holds (the emulated build's "device" memory is host memory)"""
    buf = (C.c_uint8 * max(nbytes, 1))()
    if EMU:
        C.memmove(buf, ptr, nbytes)
    else:
        from cuda.bindings import runtime as cudart
        (err,) = cudart.cudaStreamSynchronize(stream); assert int(err) == 0
        (err,) = cudart.cudaMemcpy(C.addressof(buf), ptr, nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost); assert int(err) == 0
    return bytes(buf)[:nbytes]


def dev_write(ptr, data: bytes):
    if EMU:
        C.memmove(ptr, data, len(data))
    else:
        from cuda.bindings import runtime as cudart
        (err,) = cudart.cudaMemcpy(ptr, data, len(data), cudart.cudaMemcpyKind.cudaMemcpyHostToDevice); assert int(err) == 0


def host_stage(fn):
    """A device-side stage computed on the host with the oracle's numpy recipes: reads the task list and the wave back, runs
    fn(round, tasks, rays, hits, launch_index) -> (next_rays, spawn), writes the spawned rays into their slots.  (Synchronises — a
    test harness, so that the device executor and the oracle executor see the very same rays.)"""
    def stage(wave, stream):
        n = int(np.frombuffer(dev_read(wave.d_task_count, 8, stream), np.uint64)[0])
        if n == 0:
            return
        tasks = np.frombuffer(dev_read(wave.d_tasks, 4 * n, stream), np.uint32)
        m = int(tasks.max()) + 1
        rays = np.frombuffer(dev_read(wave.d_rays, 32 * m, stream), S.RAY_DTYPE)
        hits = np.frombuffer(dev_read(wave.d_hits, 32 * m, stream), api.HIT_DTYPE)
        launch = np.frombuffer(dev_read(wave.d_launch_index, 4 * m, stream), np.uint32)
        out = fn(wave.round, tasks, rays, hits, launch)
        if out is None:
            return
        next_rays, spawn = out
        for k in np.nonzero(np.asarray(spawn, bool))[0]:
            slot = int(tasks[k])
            dev_write(wave.d_next_rays + 32 * slot, next_rays[k:k + 1].tobytes())
            dev_write(wave.d_spawn + slot, b"\x01")
    return stage


def _scene_with_two_shaded_instances():
    """two torus instances with different SBT record offsets (different closest-hit shaders) and empty space around them (miss)"""
    pos, idx = S.torus_mesh(64, 32, 1.0, 0.35)
    T, Sc, Rx, mul = S.mat4_translate, S.mat4_scale, S.mat4_rotate_x, S.mat4_mul
    sp = helpers.ScenePair((0,), True)
    b = sp.blas([(pos, idx.reshape(-1), 1)])
    inst = np.concatenate([S.make_instance(mul(mul(T(-3.0, 0, -10), Sc(3, 3, 3)), Rx(-0.5)), b, custom_index=1, sbt_offset=0),
                           S.make_instance(mul(mul(T(3.0, 0.5, -11), Sc(3, 3, 3)), Rx(0.9)), b, custom_index=2, sbt_offset=1)])
    sp.bind([sp.tlas(inst)])
    sp.build()
    return sp, (pos, idx, inst)


def test_wavefront_executor_matches_the_oracle_executor_round_by_round():
    """Three rounds over a 96 x 64 launch: ray generation, then closest-hit shader 0 (instance A) bounces every ray, closest-hit
    shader 1 (instance B) bounces only rays of even launch index, the miss shader ends the path.  Stages are computed on the host from
    the oracle's numpy recipes, so the device executor and oracle/wavefront.py run the same program on the same rays: wave sizes, task
    list sizes, task order and every hit record of every round must agree, and the task-state invariants must hold."""
    sp, (pos, idx, inst) = _scene_with_two_shaded_instances()
    W, H, ROUNDS = 96, 64, 3
    rays0 = S.pinhole_rays(W, H, 0.0, 100.0, aspect_correct=True)
    osbt = OSBT.ShaderBindingTable(1, 2, 1)
    osbt.config_hit_group(0, 0, 0, closest_hit=0); osbt.config_hit_group(0, 1, 0, closest_hit=1); osbt.config_missing(0, 0)
    sbt = sp.p.create_sbt(1, 2, 1)
    sbt.config_hit_group(0, 0, 0, api.HitGroupShaderRecord(closest_hit=0)); sbt.config_hit_group(0, 1, 0, api.HitGroupShaderRecord(closest_hit=1))
    sbt.config_missing(0, 0)
    world = [S.mat4_apply_point(i["transform"], pos) for i in inst]   # world-space vertices of each instance

    def bounce(only_even):
        def fn(rnd, tasks, rays, hits, launch):
            r, h = rays[tasks], hits[tasks]
            d = np.stack([r["dx"], r["dy"], r["dz"]], -1)
            normals = np.zeros((tasks.size, 3), f32)
            for k in range(2):
                sel = h["instance_id"] == k
                if sel.any():
                    normals[sel] = S.geometric_normals(world[k], idx, h["primitive_id"][sel], None, d[sel])
            full_r, full_h, full_n = np.zeros(tasks.size, S.RAY_DTYPE), h, normals
            full_r[:] = r
            out, src = S.bounce_rays(full_r, full_h, full_n)   # every task is a hit: one ray per task, sampled by task position
            assert src.size == tasks.size
            spawn = (launch[tasks] % 2 == 0) if only_even else np.ones(tasks.size, bool)
            return out, spawn
        return fn

    stages_closest = [bounce(False), bounce(True)]
    stages_miss = [None]
    launches = [dict(ray_flags=0x10, sbt_ray=(0, 1), miss_index=0), dict(ray_flags=0, sbt_ray=(0, 1), miss_index=0)]
    want_rows, want_waves = OW.trace_ray(sp.o, osbt, inst["sbt_offset"], W, H, lambda w, h: (rays0, np.ones(w * h, bool)), stages_closest, stages_miss,
                                         ROUNDS, launches, n_threads=4)
    seen = []

    def ray_generation(wave, stream):
        dev_write(wave.d_next_rays, rays0.tobytes())
        sp.p.stage_spawn_all(wave, stream=stream)

    def recording(stage_fn, kind, k):
        hs = host_stage(stage_fn) if stage_fn is not None else None
        def stage(wave, stream):
            n = int(np.frombuffer(dev_read(wave.d_task_count, 8, stream), np.uint64)[0])
            tasks = np.frombuffer(dev_read(wave.d_tasks, 4 * n, stream), np.uint32).copy()
            seen.append((wave.round, kind, k, tasks))
            if hs is not None:
                hs(wave, stream)
        return stage

    import torch
    st = torch.cuda.current_stream().cuda_stream
    rows = sp.p.trace_ray(sbt, W, H, ray_generation, closest_hit=[recording(f, "closest", k) for k, f in enumerate(stages_closest)],
                          miss=[recording(None, "miss", 0)], rounds=ROUNDS, round_launch=launches, stream=st, want_counts=True)
    assert rows == want_rows, (rows, want_rows)
    # task-state invariants (task-graph/src/test.rs restated): conservation per round, spawned = next wave, shrinking waves
    for r in range(1, ROUNDS + 1):
        assert rows[r]["wave"] == rows[r]["closest_tasks"] + rows[r]["miss_tasks"] + rows[r]["no_task"]
        assert rows[r]["wave"] == rows[r - 1]["spawned"]
        assert rows[r]["spawned"] <= rows[r]["closest_tasks"]
    assert rows[1]["wave"] == W * H and rows[1]["closest_tasks"] > 500 and rows[2]["wave"] > 100 and rows[1]["no_task"] == 0
    # every stage saw exactly the oracle's task list, in wave order
    for rnd, kind, k, tasks in seen:
        code = np.uint32(k) if kind == "closest" else np.uint32(k) | np.uint32(OSBT.TASK_MISS_BIT)
        assert np.array_equal(tasks, np.nonzero(want_waves[rnd - 1][3] == code)[0]), (rnd, kind, k)
    assert len(seen) == ROUNDS * 3


def test_stages_are_skipped_as_the_reference_skips_them():
    """RAY_FLAG_SKIP_CLOSEST_HIT_SHADER, an empty (u32::MAX) hit record and a missing miss shader spawn no task (trace_task.rs:206-268):
    the rays count as no_task and the pipeline is empty after one round."""
    sp, (pos, idx, inst) = _scene_with_two_shaded_instances()
    W, H = 64, 48
    rays0 = S.pinhole_rays(W, H, 0.0, 100.0, aspect_correct=True)
    import torch
    st = torch.cuda.current_stream().cuda_stream
    called = []

    def ray_generation(wave, stream):
        dev_write(wave.d_next_rays, rays0.tobytes())
        sp.p.stage_spawn_all(wave, stream=stream)

    def mark(name):
        def stage(wave, stream):
            called.append((name, int(np.frombuffer(dev_read(wave.d_task_count, 8, stream), np.uint64)[0])))
        return stage

    want = sp.o.trace(rays0, ray_flags=0x10, n_threads=4, want_counters=False)
    n_hit = int((want["instance_id"] != 0xFFFFFFFF).sum())
    n_hit_a = int((want["instance_id"] == 0).sum())
    sbt = sp.p.create_sbt(1, 2, 1)
    sbt.config_hit_group(0, 0, 0, api.HitGroupShaderRecord(closest_hit=0))   # instance B's record stays empty; no miss shader configured
    rows = sp.p.trace_ray(sbt, W, H, ray_generation, closest_hit=[mark("closest0")], miss=[mark("miss0")], rounds=2,
                          round_launch=[dict(ray_flags=0x10, sbt_ray=(0, 1))], stream=st, want_counts=True)
    assert rows[1] == dict(wave=W * H, closest_tasks=n_hit_a, miss_tasks=0, no_task=W * H - n_hit_a, spawned=0)
    assert rows[2] == dict(wave=0, closest_tasks=0, miss_tasks=0, no_task=0, spawned=0)
    assert ("closest0", n_hit_a) in called and ("miss0", 0) in called
    sbt.config_missing(0, 0)
    rows = sp.p.trace_ray(sbt, W, H, ray_generation, closest_hit=[mark("closest0")], miss=[mark("miss0")], rounds=1,
                          round_launch=[dict(ray_flags=0x10 | api.RAY_FLAG_SKIP_CLOSEST_HIT_SHADER, sbt_ray=(0, 1))], stream=st, want_counts=True)
    assert rows[1] == dict(wave=W * H, closest_tasks=0, miss_tasks=W * H - n_hit, no_task=n_hit, spawned=0)


def test_ao_frame_through_the_single_entry_equals_the_stage_by_stage_pipeline():
    """feature/ao.rs as ONE rdn_rt_trace_ray call per sample — ray generation (primary rays), round 1: closest-hit shader 0 spawns the AO
    test ray (ray type 1), round 2: closest-hit shader 1 stores "occluded", miss shader 1 stores "sky" — then the ray-gen shader's
    running mean.  The AO buffer must equal, bit for bit, the one the explicit pipeline of test_gpu_raygen.py produces (itself held to
    the oracle there), sample after sample; the counts must tell the same story."""
    import torch
    sp, _ = helpers.torus_scene(96)
    W, H = 160, 120
    n = W * H
    st = torch.cuda.current_stream().cuda_stream
    first_hit = api.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH
    sbt = sp.p.create_sbt(1, 1, 2)   # one geometry, one record offset, two ray types: primary (0) and AO (1)
    sbt.config_hit_group(0, 0, 0, api.HitGroupShaderRecord(closest_hit=0)); sbt.config_hit_group(0, 0, 1, api.HitGroupShaderRecord(closest_hit=1))
    sbt.config_missing(0, 0); sbt.config_missing(1, 1)
    d_payload = torch.ones(n, dtype=torch.float32, device="cuda")
    d_buf = torch.zeros(n, dtype=torch.float32, device="cuda")
    # the explicit pipeline
    u8 = lambda: torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    e_rays, e_hits, e_ao, e_aoh = u8(), u8(), u8(), u8()
    e_src = torch.zeros(n, dtype=torch.int32, device="cuda"); e_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    e_buf = torch.zeros(n, dtype=torch.float32, device="cuda")
    for sample in range(3):
        def ray_generation(wave, stream):
            sp.p.gen_pinhole_rays_device(wave.d_next_rays, W, H, tmin=0.0, tmax=1e30, stream=stream)
            sp.p.stage_spawn_all(wave, stream=stream)
        rows = sp.p.trace_ray(sbt, W, H, ray_generation,
                              closest_hit=[lambda wave, stream: sp.p.stage_bounce(wave, mode=1, sample_index=sample, max_sample=256, tmin=0.01, tmax=100.0, stream=stream),
                                           lambda wave, stream: sp.p.stage_store_f32(wave, 0.0, d_payload.data_ptr(), stream=stream)],
                              miss=[None, lambda wave, stream: sp.p.stage_store_f32(wave, 1.0, d_payload.data_ptr(), stream=stream)],
                              rounds=2, round_launch=[dict(ray_flags=helpers.CULL_BACK, sbt_ray=(0, 2), miss_index=0),
                                                      dict(ray_flags=first_hit, sbt_ray=(1, 2), miss_index=1)],
                              d_payload=d_payload.data_ptr(), stream=st, want_counts=True)
        sp.p.ao_resolve_device(d_payload.data_ptr(), n, sample, d_buf.data_ptr(), stream=st)
        sp.p.gen_pinhole_rays_device(e_rays.data_ptr(), W, H, tmin=0.0, tmax=1e30, stream=st)
        sp.p.trace_closest_device(e_rays.data_ptr(), n, e_hits.data_ptr(), ray_flags=helpers.CULL_BACK, grid_width=W, stream=st)
        sp.p.gen_bounce_rays_device(e_rays.data_ptr(), e_hits.data_ptr(), n, e_ao.data_ptr(), e_src.data_ptr(), e_n.data_ptr(), mode=1,
                                    sample_index=sample, max_sample=256, tmin=0.01, tmax=100.0, stream=st)
        k = int(e_n.item())
        sp.p.trace_closest_device(e_ao.data_ptr(), k, e_aoh.data_ptr(), ray_flags=first_hit, stream=st)
        sp.p.ao_accumulate_device(e_aoh.data_ptr(), e_src.data_ptr(), e_n.data_ptr(), n, sample, e_buf.data_ptr(), stream=st)
        torch.cuda.synchronize()
        occluded = int((e_aoh.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)[:k]["instance_id"] != api.INVALID_ID).sum())
        assert rows[0]["spawned"] == n and rows[1] == dict(wave=n, closest_tasks=k, miss_tasks=n - k, no_task=0, spawned=k)
        assert rows[2] == dict(wave=k, closest_tasks=occluded, miss_tasks=k - occluded, no_task=0, spawned=0)
        assert np.array_equal(d_buf.cpu().numpy(), e_buf.cpu().numpy()), sample
        assert np.all(d_payload.cpu().numpy() == 1.0)   # re-armed for the next sample
    assert 0.0 < float(d_buf.cpu().numpy().mean()) < 1.0
