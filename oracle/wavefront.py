"""TEST INFRASTRUCTURE — numpy restatement of the wavefront executor's contract (rdn_rt_trace_ray; only tests/ may import this).

What the reference does per launch (GPUWaveFrontComputeRaytracingEncoder::trace_ray, shader/ray-tracing/src/backend/wavefront_compute/
mod.rs:111-196) once the task-graph plumbing is taken away:
  * the ray generation task runs for every launch index and may call trace_ray once (mod.rs:166-174, RangedTaskSpawner :198-231);
  * `execution_round_hint` rounds (task-graph/src/runtime/mod.rs:419-461): every trace task traverses its ray
    (TraceTaskImpl::device_poll, trace_task.rs:152-204), then spawns the closest-hit task of its hit group's shader, or the miss task
    of ray.miss_index, or nothing (u32::MAX record, RAY_FLAG_SKIP_CLOSEST_HIT_SHADER) (trace_task.rs:206-268 = oracle/sbt.py dispatch);
    each spawned shader may call trace_ray again: those rays are the trace tasks of the next round;
  * every task group's alive list is compacted in order after its poll (task_group.rs:220-278), so each shader sees its tasks in wave
    order and the next wave keeps the order of the tasks that spawned it.
Task states (task_pool.rs:85-107) as they map onto this model: WAKEN = the ray is in the current wave; FINISHED = its stage spawned
nothing; GO_TO_SLEEP / SLEEP (a parent waiting for its child's payload) have no counterpart because payloads are indexed by launch
index and updated in place.  The invariants the reference's own tests assert on wake / sleep counts (task-graph/src/test.rs:1-198)
become: wave(r) = closest(r) + miss(r) + none(r); spawned(r) = wave(r + 1); a round whose stages spawn nothing empties the pipeline.

Stages are callables over numpy arrays:
    ray_generation(width, height) -> (rays[n], spawn[n] bool)
    stage(round, tasks, rays, hits, launch_index) -> (next_rays[len(tasks)], spawn[len(tasks)] bool)   or None for an empty stage
"""
from __future__ import annotations

import numpy as np

from . import sbt as SBT


def trace_ray(scene, table: SBT.ShaderBindingTable, instance_sbt_offset, width, height, ray_generation, closest_hit, miss, rounds, round_launch,
              n_threads=1):
    """Returns (count rows, waves): rows[r] = dict(wave, closest_tasks, miss_tasks, no_task, spawned); waves[r - 1] = (rays, hits,
    launch_index, task codes) of round r."""
    rays, spawn = ray_generation(width, height)
    spawn = np.asarray(spawn, bool)
    launch = np.nonzero(spawn)[0].astype(np.uint32)
    rays = rays[spawn]
    rows = [dict(wave=width * height, closest_tasks=0, miss_tasks=0, no_task=0, spawned=int(spawn.sum()))]
    waves = []
    for r in range(1, rounds + 1):
        L = dict(round_launch[min(r - 1, len(round_launch) - 1)])
        flags, sbt_ray, miss_index = L.get("ray_flags", 0), L.get("sbt_ray", (0, 0)), L.get("miss_index", 0)
        hits = scene.trace(rays, ray_flags=flags, cull_mask=L.get("cull_mask", 0xFFFFFFFF), tlas_idx=L.get("tlas_idx", 0), n_threads=n_threads,
                           want_counters=False)
        task = SBT.dispatch(table, hits, instance_sbt_offset, ray_flags=flags, sbt_ray_offset=sbt_ray[0], sbt_ray_stride=sbt_ray[1], miss_index=miss_index)
        next_rays = np.zeros_like(rays)
        spawn = np.zeros(rays.shape[0], bool)
        n_closest = n_miss = 0
        for kind, stages in (("closest", closest_hit), ("miss", miss)):
            for k, stage in enumerate(stages):
                code = np.uint32(k) if kind == "closest" else np.uint32(k) | np.uint32(SBT.TASK_MISS_BIT)
                tasks = np.nonzero(task == code)[0].astype(np.uint32)
                if kind == "closest":
                    n_closest += tasks.size
                else:
                    n_miss += tasks.size
                if stage is None or tasks.size == 0:
                    continue
                out = stage(r, tasks, rays, hits, launch)
                if out is None:
                    continue
                nr, sp = out
                sp = np.asarray(sp, bool)
                next_rays[tasks[sp]] = nr[sp]
                spawn[tasks[sp]] = True
        waves.append((rays, hits, launch, task))
        rows.append(dict(wave=int(rays.shape[0]), closest_tasks=n_closest, miss_tasks=n_miss, no_task=int(rays.shape[0]) - n_closest - n_miss,
                         spawned=int(spawn.sum())))
        rays, launch = next_rays[spawn], launch[spawn]
    return rows, waves
