"""Host-side mirror of the reference interface over the C ABI (``include/rdn_rt.h``).

Names, argument meaning and error behaviour follow rendiation:

* :class:`NaiveSahBVHSystem` — ``GPUAccelerationStructureSystemProvider`` (shader/ray-tracing/src/api/backend.rs:120-142)
  as implemented by ``NaiveSahBVHSystem`` (…/geometry/naive/mod.rs:495-610): ``create_bottom_level_acceleration_structure``,
  ``delete_…``, ``create_top_level_acceleration_structure``, ``delete_…``, ``bind_tlas``, ``bind_tlas_max_len``; plus the
  batched replacement of ``…InvocationTraversable::traverse`` (geometry/mod.rs:16-25): ``trace_closest_batch``.
* :class:`FlattenBVH`, :class:`TreeBuildOption`, ``SAH`` / ``BalanceTree`` — content/space/src/bvh, utils.rs:20-37;
  :func:`build_bvh_for_abstract_mesh`, :func:`intersect_nearest_bvh` — content/mesh/core/src/feature/bvh.rs:5-86.
* where the reference panics (= aborts, Cargo.toml:161-162) a :class:`RdnError` is raised instead.

No CPU fallback exists: loading fails loudly when ``librdn_rt.so`` is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .scenes import INSTANCE_DTYPE, RAY_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librdn_rt.so")

HIT_DTYPE = np.dtype([("t", "f4"), ("u", "f4"), ("v", "f4"), ("primitive_id", "u4"), ("geometry_id", "u4"),
                      ("instance_id", "u4"), ("instance_custom_id", "u4"), ("hit_kind", "u4")])
MESH_HIT_DTYPE = np.dtype([("px", "f4"), ("py", "f4"), ("pz", "f4"), ("distance", "f4"),
                           ("primitive_index", "u4"), ("hit", "u4"), ("pad0", "u4"), ("pad1", "u4")])
FLAT_BVH_NODE_DTYPE = np.dtype([("bmin", "f4", (3,)), ("bmax", "f4", (3,)), ("start", "u8"), ("end", "u8"), ("self_index", "u8"),
                                ("left_count", "u8"), ("has_child", "i4"), ("split_axis", "i4")])
DEV_NODE_DTYPE = np.dtype([("aabb_min", "f4", (3,)), ("hit_next", "u4"), ("aabb_max", "f4", (3,)), ("miss_next", "u4"),
                           ("range", "u4", (2,)), ("tail", "u4", (2,))])
TLAS_BOUNDING_DTYPE = np.dtype([("world_min", "f4", (3,)), ("mask", "u4"), ("world_max", "f4", (3,)), ("flags", "u4")])
INSTANCE_RECORD_DTYPE = np.dtype([("transform_inv", "f4", (16,)), ("instance_custom_index", "u4"), ("sbt_offset", "u4"),
                                  ("flags", "u4"), ("blas", "u4")])
GEOMETRY_META_DTYPE = np.dtype([("bvh_root_idx", "u4"), ("geometry_idx", "u4"), ("primitive_start", "u4"), ("geometry_flags", "u4"),
                                ("wide_root", "u4"), ("wide4_root", "u4"), ("pad", "u4", (2,))])
TRI_RECORD_DTYPE = np.dtype([("n", "f4", (3,)), ("inv_d", "f4"), ("v0", "f4", (3,)), ("uu", "f4"),
                             ("e1", "f4", (3,)), ("uv", "f4"), ("e2", "f4", (3,)), ("vv", "f4")])
WIDE_NODE_DTYPE = np.dtype([("c0_min", "f4", (3,)), ("ref0", "u4"), ("c0_max", "f4", (3,)), ("ref1", "u4"),
                            ("c1_min", "f4", (3,)), ("pad0", "u4"), ("c1_max", "f4", (3,)), ("pad1", "u4")])

ARRAYS = {  # rdn_array_id -> (name, dtype)
    0: ("tlas_binding", np.dtype("u4")), 1: ("tlas_root", np.dtype(("u4", (8,)))), 2: ("tlas_bvh_forest", DEV_NODE_DTYPE),
    3: ("tlas_bounding", TLAS_BOUNDING_DTYPE), 4: ("instances", INSTANCE_RECORD_DTYPE), 5: ("blas_meta", np.dtype(("u4", (4,)))),
    6: ("geometry_meta", GEOMETRY_META_DTYPE), 7: ("tri_bvh_forest", DEV_NODE_DTYPE), 8: ("triangles", TRI_RECORD_DTYPE),
    9: ("slot_info", np.dtype(("u4", (2,)))), 10: ("wide_nodes", WIDE_NODE_DTYPE), 11: ("prim_to_slot", np.dtype("u4")),
    12: ("irregular_instances", np.dtype("u4")),
    14: ("wide4_nodes", np.dtype([("child", [("bmin", "f4", (3,)), ("ref", "u4"), ("bmax", "f4", (3,)), ("pad", "u4")], (4,))])),
    13: ("irregular_leaf_boxes", np.dtype([("bmin", "f4", (3,)), ("pad0", "u4"), ("bmax", "f4", (3,)), ("pad1", "u4")])),
}

# RayFlagConfigRaw (api/ty.rs:102-114)
RAY_FLAG_NONE = 0x00
RAY_FLAG_FORCE_OPAQUE = 0x01
RAY_FLAG_FORCE_NON_OPAQUE = 0x02
RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH = 0x04
RAY_FLAG_SKIP_CLOSEST_HIT_SHADER = 0x08
RAY_FLAG_CULL_BACK_FACING_TRIANGLES = 0x10
RAY_FLAG_CULL_FRONT_FACING_TRIANGLES = 0x20
RAY_FLAG_CULL_OPAQUE = 0x40
RAY_FLAG_CULL_NON_OPAQUE = 0x80
RAY_FLAG_SKIP_TRIANGLES = 0x100
RAY_FLAG_SKIP_PROCEDURAL_PRIMITIVES = 0x200
# GeometryInstanceFlags / GeometryFlags / hit kinds (api/ty.rs:138-159)
GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE = 0x1
GEOMETRY_INSTANCE_TRIANGLE_FLIP_FACING = 0x2
GEOMETRY_INSTANCE_FORCE_OPAQUE = 0x4
GEOMETRY_INSTANCE_FORCE_NO_OPAQUE = 0x8
GEOMETRY_FLAG_OPAQUE = 0x1
GEOMETRY_FLAG_NO_DUPLICATE_ANYHIT_INVOCATION = 0x2
HIT_KIND_FRONT_FACING_TRIANGLE = 0xFE
HIT_KIND_BACK_FACING_TRIANGLE = 0xFF
INVALID_ID = 0xFFFFFFFF
BOUNCE_OFFSET_ORIGIN = 0x1

TRACE_AUTO, TRACE_REFERENCE_ORDER = 0, 1
TRACE_OVERLAP_PREVIOUS = 0x100  # may start while the previous device trace on the stream still walks its last rays (rdn_rt.h)
ERROR_FLAG_STACK_OVERFLOW, ERROR_FLAG_GATE_TIMEOUT = 1, 2
FACE_FRONT, FACE_BACK, FACE_DOUBLE = 0, 1, 2


class RdnError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rdn_rt error {code}: {message}")
        self.code = code


class _Launch(C.Structure):
    _fields_ = [("ray_flags", C.c_uint32), ("cull_mask", C.c_uint32), ("tlas_idx", C.c_uint32), ("grid_width", C.c_uint32),
                ("any_hit", C.c_uint32), ("sbt_ray_offset", C.c_uint32), ("sbt_ray_stride", C.c_uint32), ("miss_index", C.c_uint32)]


ANYHIT_NONE, ANYHIT_FROM_SBT = 0, 0xFFFFFFFF
ANYHIT_CONSTANT, ANYHIT_PRIMITIVE_MASK, ANYHIT_MIN_DISTANCE = 0, 1, 2
ANYHIT_BEHAVIOR_ACCEPT_HIT, ANYHIT_BEHAVIOR_END_SEARCH = 1, 2
ANYHIT_PROGRAM_DTYPE = np.dtype([("kind", "u4"), ("behavior", "u4"), ("otherwise", "u4"), ("mask", "u4"), ("value", "u4"), ("distance", "f4"),
                                 ("pad0", "u4"), ("pad1", "u4")])


def _launch(ray_flags, cull_mask, tlas_idx, grid_width, any_hit=ANYHIT_NONE, sbt_ray=(0, 0), miss_index=0):
    """rdn_launch; ``any_hit``: ANYHIT_NONE, ANYHIT_FROM_SBT, or a program index + 1 (see rdn_rt.h)"""
    return _Launch(ray_flags, cull_mask, tlas_idx, grid_width, any_hit, sbt_ray[0], sbt_ray[1], miss_index)


class _Geometry(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("n_positions", C.c_uint64), ("indices", C.c_void_p), ("n_indices", C.c_uint64),
                ("flags", C.c_uint32), ("kind", C.c_uint32)]


class _Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("bvh_visit", "bvh_hit", "tri_visit", "tri_hit", "inst_visit", "ref_abort")]


class _TraceStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("tie_rays", C.c_uint64), ("kernel_launches", C.c_uint32), ("kernel_ms", C.c_float),
                ("whole_range_rewalks", C.c_uint64)]


class _BuildStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("balance_fallbacks", "balance_fallbacks_gt10", "irregular_triangles", "irregular_instances",
                                          "reference_routed_tlas")] + \
               [(n, C.c_double) for n in ("bvh_build_ms", "flatten_ms", "upload_ms")] + [("build_threads", C.c_uint64), ("device_built_trees", C.c_uint64), ("tlas_only_commits", C.c_uint64), ("kernels_enqueued", C.c_uint64)]


class _KernelTimes(C.Structure):
    _fields_ = [("ordered_launches", C.c_uint64), ("tie_launches", C.c_uint64), ("reference_launches", C.c_uint64),
                ("ordered_ms", C.c_double), ("tie_ms", C.c_double), ("reference_ms", C.c_double)]


class _Pinhole(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rect_x", C.c_uint32), ("rect_y", C.c_uint32), ("rect_w", C.c_uint32),
                ("rect_h", C.c_uint32), ("origin", C.c_float * 3), ("tmin", C.c_float), ("tmax", C.c_float), ("aspect", C.c_float),
                ("jitter_x", C.c_float), ("jitter_y", C.c_float)]


class _Camera(C.Structure):
    _fields_ = [("view_projection_inv", C.c_float * 16), ("world_position", C.c_float * 3), ("ndc_depth", C.c_float), ("tmin", C.c_float),
                ("tmax", C.c_float), ("width", C.c_uint32), ("height", C.c_uint32), ("rect_x", C.c_uint32), ("rect_y", C.c_uint32),
                ("rect_w", C.c_uint32), ("rect_h", C.c_uint32), ("sample_index", C.c_uint32), ("pad", C.c_uint32)]


class _Bounce(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("index_base", C.c_uint32), ("scramble0", C.c_uint32), ("scramble1", C.c_uint32),
                ("sample_index", C.c_uint32), ("max_sample", C.c_uint32), ("tmin", C.c_float), ("tmax", C.c_float),
                ("flags", C.c_uint32), ("target", C.c_float * 3)]


class Wave(C.Structure):
    """rdn_wave: what a shader stage of :meth:`NaiveSahBVHSystem.trace_ray` sees (device pointers as integers)"""
    _fields_ = [("round", C.c_uint32), ("shader", C.c_uint32), ("d_tasks", C.c_void_p), ("d_task_count", C.c_void_p), ("max_tasks", C.c_uint64),
                ("d_rays", C.c_void_p), ("d_hits", C.c_void_p), ("d_launch_index", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32),
                ("d_payload", C.c_void_p), ("d_next_rays", C.c_void_p), ("d_spawn", C.c_void_p)]


STAGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Wave), C.c_void_p)


class _WaveCounts(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("wave", "closest_tasks", "miss_tasks", "no_task", "spawned")]


class _TraceRayDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("execution_round_hint", C.c_uint32), ("n_round_launch", C.c_uint32),
                ("round_launch", C.POINTER(_Launch)), ("ray_generation", STAGE_FN), ("ray_generation_user", C.c_void_p),
                ("n_closest_hit", C.c_uint32), ("n_miss", C.c_uint32), ("closest_hit", C.POINTER(STAGE_FN)), ("closest_hit_user", C.POINTER(C.c_void_p)),
                ("miss", C.POINTER(STAGE_FN)), ("miss_user", C.POINTER(C.c_void_p)), ("d_payload", C.c_void_p),
                ("counts", C.POINTER(_WaveCounts)), ("n_counts", C.c_uint32)]


class _Option(C.Structure):
    _fields_ = [("max_tree_depth", C.c_uint64), ("bin_size", C.c_uint64)]


class _PickConfig(C.Structure):
    _fields_ = [("tolerance_local", C.c_float), ("triangle_face", C.c_uint32)]


class _SbtRayConfig(C.Structure):
    _fields_ = [("ray_flags", C.c_uint32), ("sbt_ray_offset", C.c_uint32), ("sbt_ray_stride", C.c_uint32), ("miss_index", C.c_uint32)]


class _MeshView(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("n_positions", C.c_uint64), ("indices", C.c_void_p), ("n_indices", C.c_uint64)]


EXPORTED_SYMBOLS = [
    "rdn_rt_scene_create", "rdn_rt_scene_destroy", "rdn_rt_blas_create", "rdn_rt_blas_destroy", "rdn_rt_tlas_create",
    "rdn_rt_tlas_destroy", "rdn_rt_tlas_update", "rdn_rt_bind_tlas", "rdn_rt_bind_tlas_max_len", "rdn_rt_commit", "rdn_rt_trace_closest",
    "rdn_rt_trace_closest_device", "rdn_rt_trace_closest_device_n", "rdn_rt_poll_errors", "rdn_rt_host_alloc", "rdn_rt_host_free",
    "rdn_rt_host_register", "rdn_rt_host_unregister", "rdn_rt_set_any_hit_programs", "rdn_rt_bind_sbt", "rdn_rt_trace_ray",
    "rdn_rt_stage_spawn_all", "rdn_rt_stage_bounce", "rdn_rt_stage_store_f32", "rdn_rt_ao_resolve_device", "rdn_rt_trace_counted", "rdn_rt_kernel_timing_begin", "rdn_rt_kernel_timing_end",
    "rdn_rt_gen_pinhole_rays_device", "rdn_rt_gen_pinhole_rays_batch_device", "rdn_rt_gen_camera_rays_device", "rdn_rt_gen_bounce_rays_device", "rdn_rt_ao_accumulate_device", "rdn_rt_compact_u32", "rdn_rt_compact_u32_device",
    "rdn_rt_scene_blob", "rdn_rt_scene_adopt_blob", "rdn_rt_scene_array", "rdn_rt_scene_build_stats", "rdn_rt_measure_l2_read_gbs", "rdn_pick_mesh_create", "rdn_pick_mesh_destroy",
    "rdn_pick_mesh_primitive_count", "rdn_pick_mesh_nearest", "rdn_pick_mesh_all", "rdn_bvh_build", "rdn_bvh_build_device", "rdn_bvh_built_on_device", "rdn_bvh_destroy", "rdn_bvh_nodes",
    "rdn_bvh_sorted_primitive_index", "rdn_bvh_build_for_mesh", "rdn_bvh_query_nearest", "rdn_bvh_upload", "rdn_bvh_query_nearest_device", "rdn_bvh_query_list", "rdn_rt_last_error", "rdn_rt_version",
    "rdn_sbt_create", "rdn_sbt_destroy", "rdn_sbt_config_ray_generation", "rdn_sbt_config_hit_group", "rdn_sbt_config_missing", "rdn_sbt_ray_generation",
    "rdn_rt_sbt_dispatch_device", "rdn_rt_sbt_group_device", "rdn_rt_sbt_dispatch",
]

_lib = None


def lib() -> C.CDLL:
    """Load ``librdn_rt.so``; raises (no fallback) when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing — build it with `python -m rendiation_b200.build` "
                          "(rendiation_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    P = C.POINTER
    L.rdn_rt_last_error.restype = C.c_char_p
    L.rdn_rt_version.restype = C.c_char_p
    L.rdn_rt_scene_create.argtypes = [i32, P(C.c_int), P(vp)]
    L.rdn_rt_scene_destroy.argtypes = [vp]
    L.rdn_rt_scene_destroy.restype = None
    L.rdn_rt_blas_create.argtypes = [vp, P(_Geometry), u32, P(u32)]
    L.rdn_rt_blas_destroy.argtypes = [vp, u32]
    L.rdn_rt_tlas_create.argtypes = [vp, vp, u32, P(u32)]
    L.rdn_rt_tlas_destroy.argtypes = [vp, u32]
    L.rdn_rt_tlas_update.argtypes = [vp, u32, vp, u32]
    L.rdn_rt_bind_tlas.argtypes = [vp, vp, u32]
    L.rdn_rt_bind_tlas_max_len.argtypes = [vp]
    L.rdn_rt_bind_tlas_max_len.restype = u32
    L.rdn_rt_commit.argtypes = [vp]
    L.rdn_rt_trace_closest.argtypes = [vp, P(_Launch), vp, u64, vp]
    L.rdn_rt_trace_closest_device.argtypes = [vp, i32, P(_Launch), vp, u64, vp, vp, i32, P(_TraceStats)]
    L.rdn_rt_poll_errors.argtypes = [vp, i32, vp, P(u32)]
    L.rdn_rt_set_any_hit_programs.argtypes = [vp, vp, u32]
    L.rdn_rt_bind_sbt.argtypes = [vp, vp]
    L.rdn_rt_trace_ray.argtypes = [vp, i32, vp, P(_TraceRayDesc), vp]
    L.rdn_rt_stage_spawn_all.argtypes = [vp, i32, P(Wave), vp]
    L.rdn_rt_stage_bounce.argtypes = [vp, i32, P(_Bounce), P(Wave), vp]
    L.rdn_rt_stage_store_f32.argtypes = [vp, i32, P(Wave), C.c_float, vp, vp]
    L.rdn_rt_ao_resolve_device.argtypes = [vp, i32, vp, u64, u32, u32, vp, vp]
    L.rdn_rt_host_alloc.argtypes = [u64, P(vp)]
    L.rdn_rt_host_free.argtypes = [vp]
    L.rdn_rt_host_free.restype = None
    L.rdn_rt_host_register.argtypes = [vp, u64]
    L.rdn_rt_host_unregister.argtypes = [vp]
    L.rdn_rt_trace_closest_device_n.argtypes = [vp, i32, P(_Launch), vp, vp, u64, vp, vp, i32]
    L.rdn_rt_trace_counted.argtypes = [vp, P(_Launch), vp, u64, vp, P(_Counters)]
    L.rdn_rt_kernel_timing_begin.argtypes = [vp, i32]
    L.rdn_rt_kernel_timing_end.argtypes = [vp, i32, P(_KernelTimes)]
    L.rdn_rt_gen_pinhole_rays_device.argtypes = [vp, i32, P(_Pinhole), vp, vp]
    L.rdn_rt_gen_pinhole_rays_batch_device.argtypes = [vp, i32, P(_Pinhole), u32, vp, vp]
    L.rdn_rt_gen_camera_rays_device.argtypes = [vp, i32, P(_Camera), vp, vp]
    L.rdn_rt_gen_bounce_rays_device.argtypes = [vp, i32, P(_Bounce), vp, vp, u64, vp, vp, vp, vp]
    L.rdn_rt_ao_accumulate_device.argtypes = [vp, i32, vp, vp, vp, u64, u32, u32, vp, vp]
    L.rdn_rt_compact_u32.argtypes = [vp, vp, vp, u64, vp, P(u64)]
    L.rdn_rt_compact_u32_device.argtypes = [vp, i32, vp, vp, u64, vp, vp, vp]
    L.rdn_rt_scene_blob.argtypes = [vp, i32, P(vp), P(u64)]
    L.rdn_rt_scene_adopt_blob.argtypes = [vp, i32, vp, u64]
    L.rdn_rt_scene_array.argtypes = [vp, i32, vp, u64, P(u64)]
    L.rdn_rt_scene_build_stats.argtypes = [vp, P(_BuildStats)]
    L.rdn_rt_measure_l2_read_gbs.argtypes = [vp, i32, u64, i32, P(C.c_double)]
    L.rdn_pick_mesh_create.argtypes = [P(_MeshView), u32, i32, P(vp)]
    L.rdn_pick_mesh_destroy.argtypes = [vp]
    L.rdn_pick_mesh_destroy.restype = None
    L.rdn_pick_mesh_primitive_count.argtypes = [vp, P(u64)]
    L.rdn_pick_mesh_nearest.argtypes = [vp, P(_PickConfig), vp, u64, vp]
    L.rdn_pick_mesh_all.argtypes = [vp, P(_PickConfig), vp, vp, u64, P(u64)]
    L.rdn_bvh_build.argtypes = [vp, u64, i32, u32, P(_Option), P(vp)]
    L.rdn_bvh_build_device.argtypes = [vp, u64, u32, P(_Option), i32, P(vp)]
    L.rdn_bvh_built_on_device.argtypes = [vp]
    L.rdn_bvh_build_for_mesh.argtypes = [P(_MeshView), i32, u32, P(_Option), P(vp)]
    L.rdn_bvh_destroy.argtypes = [vp]
    L.rdn_bvh_destroy.restype = None
    L.rdn_bvh_nodes.argtypes = [vp, P(vp), P(u64)]
    L.rdn_bvh_sorted_primitive_index.argtypes = [vp, P(vp), P(u64)]
    L.rdn_bvh_query_nearest.argtypes = [vp, P(_MeshView), vp, u64, u32, i32, vp]
    L.rdn_bvh_upload.argtypes = [vp, P(_MeshView), i32]
    L.rdn_bvh_query_list.argtypes = [vp, P(_MeshView), vp, u64, u32, i32, vp, vp, u64, P(u64)]
    L.rdn_bvh_query_nearest_device.argtypes = [vp, vp, u64, u32, vp, vp]
    L.rdn_sbt_create.argtypes = [vp, u32, u32, u32, P(vp)]
    L.rdn_sbt_destroy.argtypes = [vp]
    L.rdn_sbt_destroy.restype = None
    L.rdn_sbt_config_ray_generation.argtypes = [vp, u32]
    L.rdn_sbt_config_hit_group.argtypes = [vp, u32, u32, u32, u32, u32, u32]
    L.rdn_sbt_config_missing.argtypes = [vp, u32, u32]
    L.rdn_sbt_ray_generation.argtypes = [vp, P(u32)]
    L.rdn_rt_sbt_dispatch_device.argtypes = [vp, i32, vp, P(_SbtRayConfig), vp, u64, vp, vp]
    L.rdn_rt_sbt_group_device.argtypes = [vp, i32, vp, vp, u64, u32, u32, vp, vp, vp]
    L.rdn_rt_sbt_dispatch.argtypes = [vp, vp, P(_SbtRayConfig), vp, u64, u32, u32, vp, vp, vp]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise RdnError(rc, lib().rdn_rt_last_error().decode("utf-8", "replace"))


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


# ------------------------------------------------------------------------------------------------ ray tracing API
@dataclass
class BlasHandle:
    id: int


@dataclass
class TlasHandle:
    id: int


@dataclass
class BottomLevelAccelerationStructureBuildSource:
    """api/backend.rs:104-118.  ``positions`` [n,3] f32; ``indices`` u32 or None; ``aabbs`` marks the AABBs variant."""
    positions: np.ndarray
    indices: np.ndarray | None = None
    flags: int = GEOMETRY_FLAG_OPAQUE
    aabbs: bool = False


class HostRegistration:
    """Page-locks a numpy array for as long as the object lives (rdn_rt_host_register / _unregister): what a Rust caller does with a
    long-lived Vec<Ray> so that the host-buffer trace runs at PCIe speed."""

    def __init__(self, array: np.ndarray):
        self.array = array  # keeps the memory alive while it is registered
        self._ptr = array.ctypes.data
        _check(lib().rdn_rt_host_register(C.c_void_p(self._ptr), array.nbytes))

    def close(self):
        if self._ptr:
            lib().rdn_rt_host_unregister(C.c_void_p(self._ptr))
            self._ptr = 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        self.close()


class NaiveSahBVHSystem:
    """The software TLAS/BLAS system with the traversal running on B200.

    ``devices``: CUDA device ordinals; the flattened scene is replicated to each and host-buffer traces are
    sharded across them (one process per GPU passes a single ordinal)."""

    def __init__(self, devices=(0,)):
        self._L = lib()
        ids = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(self._L.rdn_rt_scene_create(len(devices), ids, C.byref(h)))
        self._h = h
        self.devices = tuple(devices)

    # --- GPUAccelerationStructureSystemProvider ---
    def create_bottom_level_acceleration_structure(self, source) -> BlasHandle:
        geoms = (_Geometry * max(len(source), 1))()
        keep = []
        for g, src in zip(geoms, source):
            pos = _c(src.positions, np.float32).reshape(-1, 6 if src.aabbs else 3)
            idx = None if src.indices is None else _c(src.indices, np.uint32).reshape(-1)
            keep += [pos, idx]
            g.positions = pos.ctypes.data
            g.n_positions = pos.shape[0]
            g.indices = None if idx is None else idx.ctypes.data
            g.n_indices = 0 if idx is None else idx.size
            g.flags = src.flags
            g.kind = 1 if src.aabbs else 0
        out = C.c_uint32()
        _check(self._L.rdn_rt_blas_create(self._h, geoms, len(source), C.byref(out)))
        return BlasHandle(out.value)

    def delete_bottom_level_acceleration_structure(self, handle: BlasHandle):
        _check(self._L.rdn_rt_blas_destroy(self._h, handle.id))

    def create_top_level_acceleration_structure(self, source: np.ndarray) -> TlasHandle:
        inst = _c(source, INSTANCE_DTYPE)
        out = C.c_uint32()
        _check(self._L.rdn_rt_tlas_create(self._h, _p(inst), inst.shape[0], C.byref(out)))
        return TlasHandle(out.value)

    def update_top_level_acceleration_structure(self, handle: TlasHandle, source: np.ndarray):
        """rdn_rt_tlas_update: replace the instances of a live TLAS; the next commit rebuilds the TLAS part only."""
        inst = _c(source, INSTANCE_DTYPE)
        _check(self._L.rdn_rt_tlas_update(self._h, handle.id if isinstance(handle, TlasHandle) else int(handle), _p(inst), inst.shape[0]))

    def delete_top_level_acceleration_structure(self, handle: TlasHandle):
        _check(self._L.rdn_rt_tlas_destroy(self._h, handle.id))

    def bind_tlas(self, tlas):
        ids = _c([t.id if isinstance(t, TlasHandle) else int(t) for t in tlas], np.uint32)
        _check(self._L.rdn_rt_bind_tlas(self._h, _p(ids), ids.size))

    def bind_tlas_max_len(self) -> int:
        return int(self._L.rdn_rt_bind_tlas_max_len(self._h))

    def commit(self):
        """get_or_build_gpu_data: build + flatten + upload + replicate if anything changed."""
        _check(self._L.rdn_rt_commit(self._h))

    # --- traversal ---
    def trace_closest_batch(self, rays: np.ndarray, ray_flags: int = 0, cull_mask: int = 0xFFFFFFFF, tlas_idx: int = 0,
                            grid_width: int = 0, out: np.ndarray | None = None, **launch_kw) -> np.ndarray:
        """Host buffers in, host buffers out (H2D + traversal + D2H inside the call)."""
        rays = _c(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE) if out is None else out
        launch = _launch(ray_flags, cull_mask, tlas_idx, grid_width, **launch_kw)
        _check(self._L.rdn_rt_trace_closest(self._h, C.byref(launch), _p(rays), rays.shape[0], _p(hits)))
        return hits

    def trace_closest_host_ptr(self, rays_ptr: int, n: int, hits_ptr: int, ray_flags=0, cull_mask=0xFFFFFFFF, tlas_idx=0, grid_width=0, **launch_kw):
        """Same as :meth:`trace_closest_batch` on raw host pointers (e.g. pinned torch tensors)."""
        launch = _launch(ray_flags, cull_mask, tlas_idx, grid_width, **launch_kw)
        _check(self._L.rdn_rt_trace_closest(self._h, C.byref(launch), C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr)))

    def trace_closest_device(self, d_rays_ptr: int, n: int, d_hits_ptr: int, ray_flags=0, cull_mask=0xFFFFFFFF, tlas_idx=0,
                             grid_width=0, stream: int = 0, mode: int = TRACE_AUTO, device_index: int = 0, want_stats: bool = False,
                             overlap_previous: bool = False, **launch_kw):
        """Device-resident rays/hits; asynchronous on ``stream`` unless ``want_stats``.  ``overlap_previous``: the caller's promise
        behind RDN_TRACE_OVERLAP_PREVIOUS (nothing else enqueued on the stream since the previous trace, buffers disjoint)."""
        launch = _launch(ray_flags, cull_mask, tlas_idx, grid_width, **launch_kw)
        if overlap_previous:
            mode |= TRACE_OVERLAP_PREVIOUS
        st = _TraceStats()
        _check(self._L.rdn_rt_trace_closest_device(self._h, device_index, C.byref(launch), C.c_void_p(d_rays_ptr), n,
                                                   C.c_void_p(d_hits_ptr), C.c_void_p(stream), mode,
                                                   C.byref(st) if want_stats else None))
        if want_stats:
            return {"rays": int(st.rays), "tie_rays": int(st.tie_rays), "kernel_launches": int(st.kernel_launches),
                    "kernel_ms": float(st.kernel_ms), "whole_range_rewalks": int(st.whole_range_rewalks)}
        return None

    def trace_closest_device_n(self, d_rays_ptr: int, d_n_ptr: int, n_max: int, d_hits_ptr: int, ray_flags=0, cull_mask=0xFFFFFFFF, tlas_idx=0,
                               stream: int = 0, mode: int = TRACE_AUTO, device_index: int = 0, **launch_kw):
        """A wave whose size lives on the device (``d_n_ptr``: one u64, e.g. the count a bounce step left there); nothing
        comes back to the host."""
        launch = _launch(ray_flags, cull_mask, tlas_idx, 0, **launch_kw)
        _check(self._L.rdn_rt_trace_closest_device_n(self._h, device_index, C.byref(launch), C.c_void_p(d_rays_ptr), C.c_void_p(d_n_ptr), n_max,
                                                     C.c_void_p(d_hits_ptr), C.c_void_p(stream), mode))

    def set_any_hit_programs(self, programs):
        """The any-hit shaders of the pipeline as data: list of (kind, behavior, otherwise, mask, value, distance) — see rdn_rt.h."""
        arr = np.zeros(len(programs), ANYHIT_PROGRAM_DTYPE)
        for k, p in enumerate(programs):
            arr[k] = tuple(p) + (0, 0)
        _check(self._L.rdn_rt_set_any_hit_programs(self._h, _p(arr) if len(programs) else None, len(programs)))

    def bind_sbt(self, sbt):
        """the executor's current table (read by ANYHIT_FROM_SBT launches and trace_ray); None unbinds"""
        _check(self._L.rdn_rt_bind_sbt(self._h, sbt._h if sbt is not None else None))

    def poll_errors(self, stream: int = 0, device_index: int = 0) -> int:
        """Wait for ``stream`` and return (and clear) the safety-net flags of the asynchronous device path; raises on any."""
        flags = C.c_uint32(0)
        _check(self._L.rdn_rt_poll_errors(self._h, device_index, C.c_void_p(stream), C.byref(flags)))
        return int(flags.value)

    def trace_counted(self, rays: np.ndarray, ray_flags=0, cull_mask=0xFFFFFFFF, tlas_idx=0, **launch_kw):
        """Reference-order walk returning hits and the reference's visit counters."""
        rays = _c(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        launch = _launch(ray_flags, cull_mask, tlas_idx, 0, **launch_kw)
        ctr = _Counters()
        _check(self._L.rdn_rt_trace_counted(self._h, C.byref(launch), _p(rays), rays.shape[0], _p(hits), C.byref(ctr)))
        return hits, {n: int(getattr(ctr, n)) for n, _ in _Counters._fields_}

    def kernel_timing_begin(self, device_index: int = 0):
        """Bracket every traversal kernel launched from now on with CUDA events (measurement hook)."""
        _check(self._L.rdn_rt_kernel_timing_begin(self._h, device_index))

    def kernel_timing_end(self, device_index: int = 0) -> dict:
        """Wait for the bracketed launches; summed device milliseconds and launch counts per kernel."""
        kt = _KernelTimes()
        _check(self._L.rdn_rt_kernel_timing_end(self._h, device_index, C.byref(kt)))
        return {n: (int(getattr(kt, n)) if n.endswith("launches") else float(getattr(kt, n))) for n, _ in _KernelTimes._fields_}

    # --- ray generation / bounce on the device ---
    def gen_pinhole_rays_device(self, d_rays: int, width: int, height: int, rect=None, origin=(0.0, 0.0, 0.0), tmin=0.0, tmax=100.0,
                                aspect=1.0, jitter=(0.5, 0.5), stream: int = 0, device_index: int = 0) -> int:
        """naive/test.rs:259-264 pinhole grid (optionally a sub-rectangle, row-major) written to device rays; returns the ray count."""
        x0, y0, w, h = rect if rect is not None else (0, 0, width, height)
        p = _Pinhole(width, height, x0, y0, w, h, (C.c_float * 3)(*origin), tmin, tmax, aspect, jitter[0], jitter[1])
        _check(self._L.rdn_rt_gen_pinhole_rays_device(self._h, device_index, C.byref(p), C.c_void_p(d_rays), C.c_void_p(stream)))
        return w * h

    def gen_pinhole_rays_batch_device(self, d_rays: int, width: int, height: int, rects, jitters, origin=(0.0, 0.0, 0.0), tmin=0.0,
                                      tmax=100.0, aspect=1.0, stream: int = 0, device_index: int = 0) -> int:
        """Many rectangles (launch tiles x samples) in one launch: ``rects[k]`` with sub-pixel offset ``jitters[k]``, written
        back to back.  Returns the total ray count."""
        arr = (_Pinhole * max(len(rects), 1))()
        total = 0
        for k, ((x0, y0, w, h), (jx, jy)) in enumerate(zip(rects, jitters)):
            arr[k] = _Pinhole(width, height, x0, y0, w, h, (C.c_float * 3)(*origin), tmin, tmax, aspect, jx, jy)
            total += w * h
        _check(self._L.rdn_rt_gen_pinhole_rays_batch_device(self._h, device_index, arr, len(rects), C.c_void_p(d_rays), C.c_void_p(stream)))
        return total

    @staticmethod
    def pinhole_batch(width: int, height: int, rects, jitters, origin=(0.0, 0.0, 0.0), tmin=0.0, tmax=100.0, aspect=1.0):
        """The descriptor array of :meth:`gen_pinhole_rays_batch_device`, built once for a frame layout that is launched many
        times; returns ``(array, count, total_rays)`` for :meth:`gen_pinhole_rays_prebuilt_device`."""
        arr = (_Pinhole * max(len(rects), 1))()
        total = 0
        org = (C.c_float * 3)(*origin)
        for k, ((x0, y0, w, h), (jx, jy)) in enumerate(zip(rects, jitters)):
            arr[k] = _Pinhole(width, height, x0, y0, w, h, org, tmin, tmax, aspect, jx, jy)
            total += w * h
        return arr, len(rects), total

    def gen_pinhole_rays_prebuilt_device(self, batch, d_rays: int, stream: int = 0, device_index: int = 0) -> int:
        arr, count, total = batch
        _check(self._L.rdn_rt_gen_pinhole_rays_batch_device(self._h, device_index, arr, count, C.c_void_p(d_rays), C.c_void_p(stream)))
        return total

    def gen_camera_rays_device(self, d_rays: int, view_projection_inv, world_position, width: int, height: int, sample_index: int = 0,
                               rect=None, ndc_depth=1.0, tmin=0.0, tmax=3.4028235e38, stream: int = 0, device_index: int = 0) -> int:
        """DefaultRtxCameraInvocation::generate_ray (camera.rs:66-98) with the PCG sampler; returns the ray count."""
        x0, y0, w, h = rect if rect is not None else (0, 0, width, height)
        m = np.asarray(view_projection_inv, np.float32).reshape(-1)
        p = _Camera((C.c_float * 16)(*m), (C.c_float * 3)(*world_position), ndc_depth, tmin, tmax, width, height, x0, y0, w, h, sample_index, 0)
        _check(self._L.rdn_rt_gen_camera_rays_device(self._h, device_index, C.byref(p), C.c_void_p(d_rays), C.c_void_p(stream)))
        return w * h

    def gen_bounce_rays_device(self, d_rays_in: int, d_hits: int, n: int, d_rays_out: int, d_src_index: int, d_out_n: int, mode: int = 0,
                               index_base: int = 0, scrambles=(0x9E3779B9, 0x85EBCA6B), sample_index: int = 0, max_sample: int = 256,
                               tmin=0.01, tmax=100.0, stream: int = 0, device_index: int = 0, offset_origin: bool = False,
                               target=(0.0, 0.0, 0.0)):
        """One compacted bounce ray per primary hit (mode 0: SURVEY config-3 cosine bounce, mode 1: the AO secondary ray of
        ao.rs:249-284, mode 2: the path tracer's shadow-test ray towards a point light at ``target``, ray_hit.rs:20-45);
        ``offset_origin``: start at offset_ray_hit(hit position, geometric normal) (ray_util.rs:6-40)."""
        p = _Bounce(mode, index_base, scrambles[0], scrambles[1], sample_index, max_sample, tmin, tmax,
                    BOUNCE_OFFSET_ORIGIN if offset_origin else 0, (C.c_float * 3)(*[float(x) for x in target]))
        _check(self._L.rdn_rt_gen_bounce_rays_device(self._h, device_index, C.byref(p), C.c_void_p(d_rays_in), C.c_void_p(d_hits), n,
                                                     C.c_void_p(d_rays_out), C.c_void_p(d_src_index), C.c_void_p(d_out_n), C.c_void_p(stream)))

    def ao_accumulate_device(self, d_secondary_hits: int, d_src_index: int, d_n_secondary: int, n_pixels: int, sample_count: int,
                             d_ao_buffer: int, max_sample: int = 256, stream: int = 0, device_index: int = 0):
        """One sample of the AO frame's running mean (feature/ao.rs:187-232); see include/rdn_rt.h."""
        _check(self._L.rdn_rt_ao_accumulate_device(self._h, device_index, C.c_void_p(d_secondary_hits), C.c_void_p(d_src_index),
                                                   C.c_void_p(d_n_secondary), n_pixels, sample_count, max_sample, C.c_void_p(d_ao_buffer),
                                                   C.c_void_p(stream)))

    # --- the wavefront executor in one call (rdn_rt_trace_ray) ---
    def trace_ray(self, sbt, width: int, height: int, ray_generation, closest_hit=(), miss=(), rounds: int = 1, round_launch=(), d_payload: int = 0,
                  stream: int = 0, device_index: int = 0, want_counts: bool = False):
        """``RayTracingEncoderProvider::trace_ray``: ray generation over ``width x height``, then ``rounds`` rounds of trace -> SBT
        dispatch -> stages -> next wave, all on the device.  ``ray_generation`` / ``closest_hit[k]`` / ``miss[k]`` are Python callables
        ``f(wave: Wave, stream: int)`` that ENQUEUE work (e.g. the ``stage_*`` helpers below) and must not synchronise; ``None`` = empty
        stage.  ``round_launch``: one dict of :func:`_launch` keyword arguments (``ray_flags``, ``cull_mask``, ``tlas_idx``, ``any_hit``,
        ``sbt_ray``, ``miss_index``) per round, the last one repeating.  Returns the per-round counts when ``want_counts``."""
        errors = []

        def wrap(f):
            if f is None:
                return STAGE_FN()
            def call(_user, wave_ptr, st):
                try:
                    f(wave_ptr.contents, int(st or 0))
                    return 0
                except Exception as e:  # noqa: BLE001  (an exception must not cross the C boundary)
                    errors.append(e)
                    return -1
            return STAGE_FN(call)

        launches = [dict(rl) for rl in (round_launch or [{}])]
        L = (_Launch * len(launches))(*[_launch(rl.pop("ray_flags", 0), rl.pop("cull_mask", 0xFFFFFFFF), rl.pop("tlas_idx", 0), 0, **rl) for rl in launches])
        gen = wrap(ray_generation)
        ch = (STAGE_FN * max(len(closest_hit), 1))(*[wrap(f) for f in closest_hit])
        ms = (STAGE_FN * max(len(miss), 1))(*[wrap(f) for f in miss])
        counts = (_WaveCounts * (rounds + 1))()
        d = _TraceRayDesc(width, height, rounds, len(launches), L, gen, None, len(closest_hit), len(miss), ch, None, ms, None, d_payload,
                          counts if want_counts else None, rounds + 1 if want_counts else 0)
        rc = self._L.rdn_rt_trace_ray(self._h, device_index, sbt._h, C.byref(d), C.c_void_p(stream))
        if errors:
            raise errors[0]
        _check(rc)
        if want_counts:
            return [{n: int(getattr(c, n)) for n, _ in _WaveCounts._fields_} for c in counts]
        return None

    def stage_spawn_all(self, wave: Wave, stream: int = 0, device_index: int = 0):
        _check(self._L.rdn_rt_stage_spawn_all(self._h, device_index, C.byref(wave), C.c_void_p(stream)))

    def stage_bounce(self, wave: Wave, mode: int = 0, index_base: int = 0, scrambles=(0x9E3779B9, 0x85EBCA6B), sample_index: int = 0, max_sample: int = 256,
                     tmin=0.01, tmax=100.0, stream: int = 0, device_index: int = 0, offset_origin: bool = False, target=(0.0, 0.0, 0.0)):
        p = _Bounce(mode, index_base, scrambles[0], scrambles[1], sample_index, max_sample, tmin, tmax,
                    BOUNCE_OFFSET_ORIGIN if offset_origin else 0, (C.c_float * 3)(*[float(x) for x in target]))
        _check(self._L.rdn_rt_stage_bounce(self._h, device_index, C.byref(p), C.byref(wave), C.c_void_p(stream)))

    def stage_store_f32(self, wave: Wave, value: float, d_dst: int, stream: int = 0, device_index: int = 0):
        _check(self._L.rdn_rt_stage_store_f32(self._h, device_index, C.byref(wave), value, C.c_void_p(d_dst), C.c_void_p(stream)))

    def ao_resolve_device(self, d_payload: int, n_pixels: int, sample_count: int, d_ao_buffer: int, max_sample: int = 256, stream: int = 0, device_index: int = 0):
        _check(self._L.rdn_rt_ao_resolve_device(self._h, device_index, C.c_void_p(d_payload), n_pixels, sample_count, max_sample, C.c_void_p(d_ao_buffer),
                                                C.c_void_p(stream)))

    # --- wavefront queue compaction ---
    def compact_u32(self, values, keep):
        values = _c(values, np.uint32); keep = _c(keep, np.uint8)
        out = np.empty_like(values)
        n = C.c_uint64()
        _check(self._L.rdn_rt_compact_u32(self._h, _p(values), _p(keep), values.size, _p(out), C.byref(n)))
        return out, int(n.value)

    def compact_u32_device(self, d_in: int, d_keep: int, n: int, d_out: int, d_out_n: int, stream: int = 0, device_index: int = 0):
        _check(self._L.rdn_rt_compact_u32_device(self._h, device_index, C.c_void_p(d_in), C.c_void_p(d_keep), n, C.c_void_p(d_out),
                                                 C.c_void_p(d_out_n), C.c_void_p(stream)))

    # --- replication ---
    def blob(self, device_index: int = 0):
        ptr = C.c_void_p(); nbytes = C.c_uint64()
        _check(self._L.rdn_rt_scene_blob(self._h, device_index, C.byref(ptr), C.byref(nbytes)))
        return int(ptr.value), int(nbytes.value)

    def adopt_blob(self, d_blob_ptr: int, nbytes: int, device_index: int = 0):
        _check(self._L.rdn_rt_scene_adopt_blob(self._h, device_index, C.c_void_p(d_blob_ptr), nbytes))

    def array(self, array_id: int) -> np.ndarray:
        name, dt = ARRAYS[array_id]
        nb = C.c_uint64()
        _check(self._L.rdn_rt_scene_array(self._h, array_id, None, 0, C.byref(nb)))
        out = np.zeros(int(nb.value) // dt.itemsize, dt)
        if nb.value:
            _check(self._L.rdn_rt_scene_array(self._h, array_id, _p(out), nb.value, C.byref(nb)))
        return out

    def measure_l2_read_gbs(self, nbytes: int = 64 << 20, passes: int = 50, device_index: int = 0) -> float:
        """read bandwidth (GB/s) of an L2-resident buffer: the roofline's L2 denominator (measurement hook)"""
        out = C.c_double()
        _check(self._L.rdn_rt_measure_l2_read_gbs(self._h, device_index, nbytes, passes, C.byref(out)))
        return float(out.value)

    def build_stats(self) -> dict:
        """What the flattener found: SAH->BalanceTree fallbacks and the irregular triangles / instances (include/rdn_rt.h)."""
        st = _BuildStats()
        _check(self._L.rdn_rt_scene_build_stats(self._h, C.byref(st)))
        return {n: (float(getattr(st, n)) if t is C.c_double else int(getattr(st, n))) for n, t in _BuildStats._fields_}

    def arrays(self) -> dict:
        return {ARRAYS[i][0]: self.array(i) for i in ARRAYS}

    def create_sbt(self, max_geometry_count_in_blas: int, max_tlas_offset: int, ray_type_count: int) -> "ShaderBindingTable":
        """``GPURayTracingDeviceProvider::create_sbt`` (api/backend.rs:74-79)"""
        return ShaderBindingTable(self, max_geometry_count_in_blas, max_tlas_offset, ray_type_count)

    def close(self):
        if getattr(self, "_h", None):
            self._L.rdn_rt_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ shader binding table
SBT_NO_SHADER = 0xFFFFFFFF
TASK_NONE = 0xFFFFFFFF
TASK_MISS_BIT = 0x80000000
RAY_FLAG_SKIP_CLOSEST_HIT_SHADER = 0x08


class HitGroupShaderRecord:
    """``HitGroupShaderRecord`` (shader/ray-tracing/src/api/backend.rs:83-88); ``None`` = no shader."""

    def __init__(self, closest_hit=None, any_hit=None, intersection=None):
        self.closest_hit, self.any_hit, self.intersection = closest_hit, any_hit, intersection


class ShaderBindingTable:
    """``ShaderBindingTableProvider`` (api/backend.rs:90-101) as created by ``GPURayTracingDeviceProvider::create_sbt``
    (api/backend.rs:74-79), plus the dispatch the wavefront executor performs with it after traversal
    (wavefront_compute/trace_task.rs:206-268): a task code per ray and the per-shader task lists."""

    def __init__(self, system: "NaiveSahBVHSystem", max_geometry_count_in_blas: int, max_tlas_offset: int, ray_type_count: int):
        self._L = lib()
        self._sys = system
        h = C.c_void_p()
        _check(self._L.rdn_sbt_create(system._h, max_geometry_count_in_blas, max_tlas_offset, ray_type_count, C.byref(h)))
        self._h = h

    @staticmethod
    def _handle(s) -> int:
        return SBT_NO_SHADER if s is None else int(s)

    def config_ray_generation(self, s: int):
        _check(self._L.rdn_sbt_config_ray_generation(self._h, int(s)))

    def config_hit_group(self, geometry_idx: int, tlas_offset: int, ray_ty_idx: int, hit_group: HitGroupShaderRecord):
        _check(self._L.rdn_sbt_config_hit_group(self._h, geometry_idx, tlas_offset, ray_ty_idx, self._handle(hit_group.closest_hit),
                                                self._handle(hit_group.any_hit), self._handle(hit_group.intersection)))

    def config_missing(self, ray_ty_idx: int, s: int):
        _check(self._L.rdn_sbt_config_missing(self._h, ray_ty_idx, int(s)))

    @property
    def ray_generation(self) -> int:
        v = C.c_uint32()
        _check(self._L.rdn_sbt_ray_generation(self._h, C.byref(v)))
        return int(v.value)

    def dispatch(self, hits: np.ndarray, n_closest_shaders: int, n_miss_shaders: int, ray_flags: int = 0, sbt_ray_offset: int = 0,
                 sbt_ray_stride: int = 1, miss_index: int = 0):
        """host buffers: returns (task code per ray, ray indices grouped by shader, group offsets)"""
        hits = _c(hits, HIT_DTYPE)
        n = hits.shape[0]
        task = np.zeros(n, np.uint32)
        queue = np.zeros(max(n, 1), np.uint32)
        offsets = np.zeros(n_closest_shaders + n_miss_shaders + 1, np.uint64)
        cfg = _SbtRayConfig(ray_flags, sbt_ray_offset, sbt_ray_stride, miss_index)
        _check(self._L.rdn_rt_sbt_dispatch(self._sys._h, self._h, C.byref(cfg), _p(hits), n, n_closest_shaders, n_miss_shaders, _p(task),
                                           _p(queue), _p(offsets)))
        return task, queue[:int(offsets[-1])], offsets

    def dispatch_device(self, d_hits: int, n: int, d_task: int, ray_flags: int = 0, sbt_ray_offset: int = 0, sbt_ray_stride: int = 1,
                        miss_index: int = 0, stream: int = 0, device_index: int = 0):
        cfg = _SbtRayConfig(ray_flags, sbt_ray_offset, sbt_ray_stride, miss_index)
        _check(self._L.rdn_rt_sbt_dispatch_device(self._sys._h, device_index, self._h, C.byref(cfg), d_hits, n, d_task, stream))

    def group_device(self, d_task: int, n: int, n_closest_shaders: int, n_miss_shaders: int, d_queue: int, d_offsets: int, stream: int = 0,
                     device_index: int = 0):
        _check(self._L.rdn_rt_sbt_group_device(self._sys._h, device_index, self._h, d_task, n, n_closest_shaders, n_miss_shaders, d_queue,
                                               d_offsets, stream))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            self._L.rdn_sbt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ space query API
@dataclass
class TreeBuildOption:
    """content/space/src/utils.rs:20-32 (defaults 10 / 50)."""
    max_tree_depth: int = 10
    bin_size: int = 50


@dataclass
class SAH:
    """content/space/src/bvh/strategy.rs:99-117: ``SAH::new(pre_partition_check_count)``."""
    pre_partition_check_count: int = 4


class BalanceTree:
    """content/space/src/bvh/strategy.rs:65."""


# MeshPrimitiveTopology (content/mesh/core/src/primitive.rs:108-128)
TOPOLOGY_POINT_LIST, TOPOLOGY_LINE_LIST, TOPOLOGY_LINE_STRIP, TOPOLOGY_TRIANGLE_LIST, TOPOLOGY_TRIANGLE_STRIP = range(5)


class PickMesh:
    """A device-resident attribute mesh for brute-force picking: ``ray_intersect_nearest`` / ``ray_intersect_all``
    (content/mesh/core/src/feature/intersection.rs:3-37) with ``MeshBufferIntersectConfig`` (container/attributes/picking.rs:4-27)."""

    def __init__(self, positions, indices=None, topology: int = TOPOLOGY_TRIANGLE_LIST, device: int = 0):
        self._L = lib()
        self._pos = _c(positions, np.float32).reshape(-1, 3)
        self._idx = None if indices is None else _c(indices, np.uint32).reshape(-1)
        view = _MeshView(self._pos.ctypes.data, self._pos.shape[0], None if self._idx is None else self._idx.ctypes.data,
                         0 if self._idx is None else self._idx.size)
        h = C.c_void_p()
        _check(self._L.rdn_pick_mesh_create(C.byref(view), topology, device, C.byref(h)))
        self._h = h

    @property
    def primitive_count(self) -> int:
        n = C.c_uint64()
        _check(self._L.rdn_pick_mesh_primitive_count(self._h, C.byref(n)))
        return int(n.value)

    def ray_intersect_nearest(self, rays, tolerance_local: float = 0.0, triangle_face: int = FACE_DOUBLE) -> np.ndarray:
        rays = _c(rays, RAY_DTYPE)
        out = np.zeros(rays.shape[0], MESH_HIT_DTYPE)
        cfg = _PickConfig(tolerance_local, triangle_face)
        _check(self._L.rdn_pick_mesh_nearest(self._h, C.byref(cfg), _p(rays), rays.shape[0], _p(out)))
        return out

    def ray_intersect_all(self, ray, tolerance_local: float = 0.0, triangle_face: int = FACE_DOUBLE) -> np.ndarray:
        ray = _c(np.asarray(ray).reshape(1), RAY_DTYPE)
        cfg = _PickConfig(tolerance_local, triangle_face)
        total = C.c_uint64()
        _check(self._L.rdn_pick_mesh_all(self._h, C.byref(cfg), _p(ray), None, 0, C.byref(total)))
        out = np.zeros(int(total.value), MESH_HIT_DTYPE)
        if total.value:
            _check(self._L.rdn_pick_mesh_all(self._h, C.byref(cfg), _p(ray), _p(out), out.shape[0], C.byref(total)))
        return out

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            self._L.rdn_pick_mesh_destroy(self._h)
            self._h = None


class FlattenBVH:
    """content/space/src/bvh/mod.rs:26-79: ``FlattenBVH::new(boxes, &mut strategy, &option)``; boxes = [n,6] (min,max)."""

    def __init__(self, boxes=None, strategy=None, option: TreeBuildOption | None = None, _handle=None, device: int | None = None):
        """``device``: build on that CUDA device (SAH only; same tree — what the device build does not cover falls to the host builder,
        see ``built_on_device``)"""
        self._L = lib()
        if _handle is not None:
            self._h = _handle
            return
        boxes = _c(boxes, np.float32).reshape(-1, 6)
        strategy = strategy if strategy is not None else SAH(4)
        option = option or TreeBuildOption()
        opt = _Option(option.max_tree_depth, option.bin_size)
        h = C.c_void_p()
        if device is not None and isinstance(strategy, SAH):
            rc = self._L.rdn_bvh_build_device(_p(boxes), boxes.shape[0], strategy.pre_partition_check_count, C.byref(opt), device, C.byref(h))
        elif isinstance(strategy, SAH):
            rc = self._L.rdn_bvh_build(_p(boxes), boxes.shape[0], 0, strategy.pre_partition_check_count, C.byref(opt), C.byref(h))
        else:
            rc = self._L.rdn_bvh_build(_p(boxes), boxes.shape[0], 1, 0, C.byref(opt), C.byref(h))
        _check(rc)
        self._h = h

    @property
    def built_on_device(self) -> bool:
        return bool(self._L.rdn_bvh_built_on_device(self._h))

    @property
    def nodes(self) -> np.ndarray:
        ptr = C.c_void_p(); n = C.c_uint64()
        _check(self._L.rdn_bvh_nodes(self._h, C.byref(ptr), C.byref(n)))
        buf = (C.c_char * (n.value * FLAT_BVH_NODE_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=FLAT_BVH_NODE_DTYPE).copy()

    @property
    def sorted_primitive_index(self) -> np.ndarray:
        ptr = C.c_void_p(); n = C.c_uint64()
        _check(self._L.rdn_bvh_sorted_primitive_index(self._h, C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, np.uint64)
        buf = (C.c_char * (n.value * 8)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint64).copy()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.rdn_bvh_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _mesh_view(positions, indices):
    pos = _c(positions, np.float32).reshape(-1, 3)
    idx = _c(indices, np.uint32).reshape(-1)
    return _MeshView(pos.ctypes.data, pos.shape[0], idx.ctypes.data, idx.size), (pos, idx)


def build_bvh_for_abstract_mesh(positions, indices, strategy=None, option: TreeBuildOption | None = None) -> FlattenBVH:
    """content/mesh/core/src/feature/bvh.rs:5-21 over an indexed triangle list."""
    L = lib()
    mv, _keep = _mesh_view(positions, indices)
    strategy = strategy if strategy is not None else SAH(4)
    option = option or TreeBuildOption()
    opt = _Option(option.max_tree_depth, option.bin_size)
    h = C.c_void_p()
    if isinstance(strategy, SAH):
        rc = L.rdn_bvh_build_for_mesh(C.byref(mv), 0, strategy.pre_partition_check_count, C.byref(opt), C.byref(h))
    else:
        rc = L.rdn_bvh_build_for_mesh(C.byref(mv), 1, 0, C.byref(opt), C.byref(h))
    _check(rc)
    return FlattenBVH(_handle=h)


def intersect_list_bvh(positions, indices, rays, bvh: FlattenBVH, face_side: int = FACE_DOUBLE, device: int = 0):
    """content/mesh/core/src/feature/bvh.rs:23-55 for a batch of rays: ``(offsets[n+1], hits[total])`` — the hits of ray ``i`` are
    ``hits[offsets[i]:offsets[i+1]]`` in the reference's visiting order."""
    L = lib()
    mv, _keep = _mesh_view(positions, indices)
    rays = _c(rays, RAY_DTYPE)
    offsets = np.zeros(rays.shape[0] + 1, np.uint64)
    total = C.c_uint64()
    _check(L.rdn_bvh_query_list(bvh._h, C.byref(mv), _p(rays), rays.shape[0], face_side, device, _p(offsets), None, 0, C.byref(total)))
    out = np.zeros(int(total.value), MESH_HIT_DTYPE)
    if total.value:
        _check(L.rdn_bvh_query_list(bvh._h, C.byref(mv), _p(rays), rays.shape[0], face_side, device, _p(offsets), _p(out), int(total.value), C.byref(total)))
    return offsets, out


def upload_bvh(bvh: FlattenBVH, positions, indices, device: int = 0) -> None:
    """Make the space query device-resident: nodes + the mesh gathered in BVH order are uploaded once."""
    mv, _keep = _mesh_view(positions, indices)
    _check(lib().rdn_bvh_upload(bvh._h, C.byref(mv), device))


def intersect_nearest_bvh_device(bvh: FlattenBVH, d_rays: int, n: int, d_out: int, face_side: int = FACE_DOUBLE, stream: int = 0) -> None:
    """intersect_nearest_bvh on device buffers (after :func:`upload_bvh`); asynchronous on ``stream``."""
    _check(lib().rdn_bvh_query_nearest_device(bvh._h, C.c_void_p(d_rays), n, face_side, C.c_void_p(d_out), C.c_void_p(stream)))


def intersect_nearest_bvh(positions, indices, rays, bvh: FlattenBVH, face_side: int = FACE_DOUBLE, device: int = 0) -> np.ndarray:
    """content/mesh/core/src/feature/bvh.rs:57-86 for a batch of rays; ``hit == 0`` is ``OptionalNearest::none()``."""
    L = lib()
    mv, _keep = _mesh_view(positions, indices)
    rays = _c(rays, RAY_DTYPE)
    out = np.zeros(rays.shape[0], MESH_HIT_DTYPE)
    _check(L.rdn_bvh_query_nearest(bvh._h, C.byref(mv), _p(rays), rays.shape[0], face_side, device, _p(out)))
    return out
