"""TEST INFRASTRUCTURE — numpy restatement of the reference's shader-binding-table dispatch (the decision the wavefront executor
takes per ray after traversal).  Only tests/ may import this.

Follows
  ShaderBindingTableInfo::config_hit_group / config_missing / config_ray_generation
      shader/ray-tracing/src/backend/wavefront_compute/sbt.rs:13-50   (record index = ray_ty_idx + geometry_idx * ray_stride + tlas_offset)
  HitCtxInfoVar::compute_sbt_hit_group       shader/ray-tracing/src/api/ctx.rs:53-55    (offset + stride * geometry_id + instance_sbt_offset)
  get_closest_handle / get_missing_handle    wavefront_compute/sbt.rs:252-268
  TraceTaskImpl::device_poll                 wavefront_compute/trace_task.rs:206-268    (hit -> closest task unless SKIP_CLOSEST_HIT_SHADER or
                                                                                          u32::MAX; miss -> miss task unless u32::MAX)
Parity unpinned upstream: the reference has no test of this selection; the table layout and the formula are restated as written.
"""
from __future__ import annotations

import numpy as np

NO_SHADER = 0xFFFFFFFF
TASK_NONE = 0xFFFFFFFF
TASK_MISS_BIT = 0x80000000
RAY_FLAG_SKIP_CLOSEST_HIT_SHADER = 0x08
INVALID = 0xFFFFFFFF


class ShaderBindingTable:
    def __init__(self, max_geometry_count_in_blas: int, max_tlas_offset: int, ray_type_count: int):
        self.ray_stride = ray_type_count
        n = max_geometry_count_in_blas * max_tlas_offset * ray_type_count
        self.ray_hit = np.full((n, 3), NO_SHADER, np.uint32)  # closest_hit, any_hit, intersection
        self.ray_miss = np.full(ray_type_count, NO_SHADER, np.uint32)
        self.ray_gen = NO_SHADER

    def config_ray_generation(self, s):
        self.ray_gen = s

    def config_hit_group(self, geometry_idx, tlas_offset, ray_ty_idx, closest_hit=None, any_hit=None, intersection=None):
        idx = ray_ty_idx + geometry_idx * self.ray_stride + tlas_offset
        if idx >= self.ray_hit.shape[0]:
            raise IndexError("the reference's set_value(..).unwrap() panics here")
        self.ray_hit[idx] = [NO_SHADER if v is None else v for v in (closest_hit, any_hit, intersection)]

    def config_missing(self, ray_ty_idx, s):
        self.ray_miss[ray_ty_idx] = s


def dispatch(sbt: ShaderBindingTable, hits: np.ndarray, instance_sbt_offset: np.ndarray, ray_flags=0, sbt_ray_offset=0, sbt_ray_stride=1,
             miss_index=0) -> np.ndarray:
    """task code per ray; instance_sbt_offset[k] = instance_shader_binding_table_record_offset of TLAS slot k"""
    n = hits.shape[0]
    task = np.full(n, TASK_NONE, np.uint32)
    is_hit = hits["instance_id"] != INVALID
    if not (ray_flags & RAY_FLAG_SKIP_CLOSEST_HIT_SHADER):
        idx = np.nonzero(is_hit)[0]
        group = (np.uint64(sbt_ray_offset) + np.uint64(sbt_ray_stride) * hits["geometry_id"][idx].astype(np.uint64)
                 + instance_sbt_offset[hits["instance_id"][idx]].astype(np.uint64)) & np.uint64(0xFFFFFFFF)   # u32 arithmetic wraps
        ok = group < sbt.ray_hit.shape[0]
        shader = np.full(idx.shape[0], NO_SHADER, np.uint32)
        shader[ok] = sbt.ray_hit[group[ok].astype(np.int64), 0]
        task[idx] = np.where(shader == NO_SHADER, TASK_NONE, shader)
    if miss_index < sbt.ray_miss.shape[0] and sbt.ray_miss[miss_index] != NO_SHADER:
        task[~is_hit] = np.uint32(sbt.ray_miss[miss_index]) | np.uint32(TASK_MISS_BIT)
    return task


def group(task: np.ndarray, n_closest_shaders: int, n_miss_shaders: int):
    """ray indices grouped by task (closest shaders first, then miss shaders), ray order inside a group"""
    queue, offsets = [], [0]
    for b in range(n_closest_shaders + n_miss_shaders):
        code = b if b < n_closest_shaders else ((b - n_closest_shaders) | TASK_MISS_BIT)
        queue.append(np.nonzero(task == np.uint32(code))[0].astype(np.uint32))
        offsets.append(offsets[-1] + queue[-1].shape[0])
    return (np.concatenate(queue) if queue else np.zeros(0, np.uint32)), np.asarray(offsets, np.uint64)
