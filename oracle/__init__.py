"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for ``oracle/_build/liboracle.so`` (the plain-C CPU restatement of rendiation's
BVH closest-hit path, see ``oracle/oracle.h``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this package; the product
(``rendiation_b200``) never does.

Parity pinning: the reference has no golden vectors for traversal results ("parity unpinned",
see ``oracle/oracle.h`` header); the KATs it does hold are replayed in ``tests/test_oracle_kat.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

RAY_DTYPE = np.dtype([("ox", "f4"), ("oy", "f4"), ("oz", "f4"), ("tmin", "f4"),
                      ("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("tmax", "f4")])
HIT_DTYPE = np.dtype([("t", "f4"), ("u", "f4"), ("v", "f4"), ("primitive_id", "u4"), ("geometry_id", "u4"),
                      ("instance_id", "u4"), ("instance_custom_id", "u4"), ("hit_kind", "u4")])
MESH_HIT_DTYPE = np.dtype([("px", "f4"), ("py", "f4"), ("pz", "f4"), ("distance", "f4"),
                           ("primitive_index", "u4"), ("hit", "u4"), ("pad0", "u4"), ("pad1", "u4")])
INSTANCE_DTYPE = np.dtype([("transform", "f4", (16,)), ("instance_custom_index", "u4"), ("mask", "u4"),
                           ("sbt_offset", "u4"), ("flags", "u4"), ("blas_handle", "u4")])
DEV_NODE_DTYPE = np.dtype([("aabb_min", "f4", (3,)), ("hit_next", "u4"), ("aabb_max", "f4", (3,)), ("miss_next", "u4"),
                           ("range", "u4", (2,)), ("tail", "u4", (2,))])
DEV_INSTANCE_DTYPE = np.dtype([("transform", "f4", (16,)), ("transform_inv", "f4", (16,)),
                               ("instance_custom_index", "u4"), ("sbt_offset", "u4"), ("flags", "u4"), ("blas", "u4")])
TLAS_BOUNDING_DTYPE = np.dtype([("world_min", "f4", (3,)), ("mask", "u4"), ("world_max", "f4", (3,)), ("flags", "u4")])
GEOM_META_DTYPE = np.dtype([("bvh_root_idx", "u4"), ("geometry_idx", "u4"), ("primitive_start", "u4"), ("geometry_flags", "u4")])
BVH_NODE_DTYPE = np.dtype([("bmin", "f4", (3,)), ("bmax", "f4", (3,)), ("start", "u8"), ("end", "u8"), ("self_index", "u8"),
                           ("left_count", "u8"), ("has_child", "i4"), ("split_axis", "i4")])

STRATEGY_SAH, STRATEGY_BALANCE = 0, 1
FACE_FRONT, FACE_BACK, FACE_DOUBLE = 0, 1, 2


CANDIDATE_DTYPE = np.dtype([("distance", "f4"), ("t_object", "f4"), ("u", "f4"), ("v", "f4"), ("sign", "f4"), ("scaling", "f4"),
                            ("instance_id", "u4"), ("geometry_id", "u4"), ("primitive_id", "u4"), ("slot", "u4"), ("in_range", "u4"), ("pad", "u4")])


ANYHIT_PROGRAM_DTYPE = np.dtype([("kind", "u4"), ("behavior", "u4"), ("otherwise", "u4"), ("mask", "u4"), ("value", "u4"), ("distance", "f4"),
                                 ("pad0", "u4"), ("pad1", "u4")])
ANYHIT_CONSTANT, ANYHIT_PRIMITIVE_MASK, ANYHIT_MIN_DISTANCE = 0, 1, 2
ANYHIT_ACCEPT, ANYHIT_END_SEARCH = 1, 2


class _AnyHitSetup(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("uniform_program", C.c_uint32), ("programs", C.c_void_p), ("n_programs", C.c_uint32),
                ("hit_group_any", C.c_void_p), ("n_hit_groups", C.c_uint32), ("sbt_ray_offset", C.c_uint32), ("sbt_ray_stride", C.c_uint32)]


class Launch(C.Structure):
    _fields_ = [("ray_flags", C.c_uint32), ("cull_mask", C.c_uint32), ("tlas_idx", C.c_uint32), ("grid_width", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("bvh_visit", "bvh_hit", "tri_visit", "tri_hit", "inst_visit", "ref_abort")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class _Bvh(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("n_nodes", C.c_uint64), ("cap_nodes", C.c_uint64),
                ("sorted_primitive_index", C.POINTER(C.c_uint64)), ("n_prims", C.c_uint64),
                ("balance_fallbacks", C.c_uint64), ("balance_fallbacks_gt10", C.c_uint64), ("error", C.c_int32)]


class _SceneView(C.Structure):
    _names = ["tlas_binding", "tlas_bvh_root", "tlas_bvh_forest", "tlas_data", "tlas_bounding", "blas_meta_info",
              "tri_bvh_root", "tri_bvh_forest", "indices_redirect", "indices", "vertices"]
    _fields_ = sum(([(n, C.c_void_p), ("n_" + n, C.c_uint64)] for n in _names), []) + \
        [("balance_fallbacks", C.c_uint64), ("balance_fallbacks_gt10", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``make -C oracle``)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.orc_bvh_build.restype = C.POINTER(_Bvh)
    L.orc_bvh_build.argtypes = [vp, u64, i32, u32, u64, u64]
    L.orc_bvh_free.argtypes = [C.POINTER(_Bvh)]
    L.orc_bvh_compute_next.argtypes = [C.POINTER(_Bvh), vp]
    L.orc_ray_box_a.restype = i32
    L.orc_ray_box_a.argtypes = [vp, vp]
    L.orc_patha_query_nearest.argtypes = [C.POINTER(_Bvh), vp, vp, vp, u64, i32, vp, i32]
    L.orc_patha_query_list.argtypes = [C.POINTER(_Bvh), vp, vp, vp, u64, i32, vp, vp, u64]
    L.orc_patha_query_list.restype = u64
    L.orc_brute_query_nearest.argtypes = [vp, vp, u64, vp, u64, i32, vp, i32]
    L.orc_scene_new.restype = vp
    L.orc_scene_free.argtypes = [vp]
    L.orc_scene_create_blas.restype = u32
    L.orc_scene_create_blas.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp]
    L.orc_scene_delete_blas.argtypes = [vp, u32]
    L.orc_scene_create_tlas.restype = u32
    L.orc_scene_create_tlas.argtypes = [vp, vp, u32]
    L.orc_scene_delete_tlas.argtypes = [vp, u32]
    L.orc_scene_bind_tlas.argtypes = [vp, vp, u32]
    L.orc_scene_build.restype = i32
    L.orc_scene_build.argtypes = [vp]
    L.orc_scene_get_view.argtypes = [vp, C.POINTER(_SceneView)]
    L.orc_scene_trace.restype = i32
    L.orc_scene_trace.argtypes = [vp, C.POINTER(Launch), vp, u64, vp, C.POINTER(Counters), i32]
    L.orc_workgroup_inclusive_scan_u32.argtypes = [vp, u64, u32, vp]
    L.orc_inclusive_scan_u32.argtypes = [vp, u64, vp]
    L.orc_stream_compaction_u32.restype = u64
    L.orc_stream_compaction_u32.argtypes = [vp, vp, u64, vp]
    L.orc_shuffle_move_u32.argtypes = [vp, vp, vp, u64, vp]
    L.orc_mat4_compose.argtypes = [vp, vp, vp]
    L.orc_mat4_inverse_or_identity.argtypes = [vp, vp]
    L.orc_mat4_mul_vec4.argtypes = [vp, vp, vp]
    _lib = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class FlattenBVH:
    """content/space FlattenBVH built by the oracle (``FlattenBVH::new``)."""

    def __init__(self, boxes: np.ndarray, strategy: int = STRATEGY_SAH, sah_buckets: int = 4,
                 max_tree_depth: int = 10, bin_size: int = 50):
        boxes = _c(boxes, np.float32).reshape(-1, 6)
        self._h = lib().orc_bvh_build(_p(boxes), boxes.shape[0], strategy, sah_buckets, max_tree_depth, bin_size)
        b = self._h.contents
        self.n_nodes = int(b.n_nodes)
        self.n_prims = int(b.n_prims)
        self.balance_fallbacks = int(b.balance_fallbacks)
        self.balance_fallbacks_gt10 = int(b.balance_fallbacks_gt10)
        self.error = int(b.error)

    @property
    def nodes(self) -> np.ndarray:
        b = self._h.contents
        buf = (C.c_char * (self.n_nodes * BVH_NODE_DTYPE.itemsize)).from_address(b.nodes)
        return np.frombuffer(buf, dtype=BVH_NODE_DTYPE).copy()

    @property
    def sorted_primitive_index(self) -> np.ndarray:
        b = self._h.contents
        return np.ctypeslib.as_array(b.sorted_primitive_index, shape=(max(self.n_prims, 1),))[:self.n_prims].copy()

    def compute_next(self) -> np.ndarray:
        out = np.zeros((self.n_nodes, 2), np.uint32)
        lib().orc_bvh_compute_next(self._h, _p(out))
        return out

    def query_nearest(self, positions, indices, rays, face_side=FACE_DOUBLE, n_threads=1) -> np.ndarray:
        positions = _c(positions, np.float32); indices = _c(indices, np.uint32)
        rays = _c(rays, RAY_DTYPE)
        out = np.zeros(rays.shape[0], MESH_HIT_DTYPE)
        lib().orc_patha_query_nearest(self._h, _p(positions), _p(indices), _p(rays), rays.shape[0], face_side, _p(out), n_threads)
        return out

    def query_list(self, positions, indices, rays, face_side=FACE_DOUBLE):
        """intersect_list_bvh for every ray: (offsets[n+1], hits[total]) in the reference's visiting order"""
        positions = _c(positions, np.float32); indices = _c(indices, np.uint32)
        rays = _c(rays, RAY_DTYPE)
        offsets = np.zeros(rays.shape[0] + 1, np.uint64)
        L = lib()
        total = L.orc_patha_query_list(self._h, _p(positions), _p(indices), _p(rays), rays.shape[0], face_side, _p(offsets), None, 0)
        out = np.zeros(int(total), MESH_HIT_DTYPE)
        L.orc_patha_query_list(self._h, _p(positions), _p(indices), _p(rays), rays.shape[0], face_side, _p(offsets), _p(out), int(total))
        return offsets, out

    def __del__(self):
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.orc_bvh_free(self._h)
            self._h = None


def brute_query_nearest(positions, indices, rays, face_side=FACE_DOUBLE, n_threads=1) -> np.ndarray:
    positions = _c(positions, np.float32); indices = _c(indices, np.uint32); rays = _c(rays, RAY_DTYPE)
    out = np.zeros(rays.shape[0], MESH_HIT_DTYPE)
    lib().orc_brute_query_nearest(_p(positions), _p(indices), indices.size // 3, _p(rays), rays.shape[0], face_side, _p(out), n_threads)
    return out


def ray_box_a(ray6, box6) -> bool:
    r = _c(ray6, np.float32); b = _c(box6, np.float32)
    return bool(lib().orc_ray_box_a(_p(r), _p(b)))


class Scene:
    """NaiveSahBVHSystem restated: create/delete BLAS & TLAS, bind, build, traverse."""

    def __init__(self):
        self._h = lib().orc_scene_new()

    def create_blas(self, geometries) -> int:
        """geometries: list of (positions[n,3] f32, indices u32 | None, flags) or (.., is_aabb=True)"""
        n = len(geometries)
        pos = [_c(g[0], np.float32).reshape(-1, 3) for g in geometries]
        idx = [None if g[1] is None else _c(g[1], np.uint32).reshape(-1) for g in geometries]
        flags = np.array([g[2] for g in geometries], np.uint32)
        aabb = np.array([1 if (len(g) > 3 and g[3]) else 0 for g in geometries], np.uint8)
        pos_ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pos])
        idx_ptrs = (C.c_void_p * n)(*[None if i is None else i.ctypes.data for i in idx])
        n_pos = np.array([p.shape[0] for p in pos], np.uint64)
        n_idx = np.array([0 if i is None else i.size for i in idx], np.uint64)
        return int(lib().orc_scene_create_blas(self._h, n, pos_ptrs, _p(n_pos), idx_ptrs, _p(n_idx), _p(flags), _p(aabb)))

    def delete_blas(self, h: int):
        lib().orc_scene_delete_blas(self._h, h)

    def create_tlas(self, instances: np.ndarray) -> int:
        inst = _c(instances, INSTANCE_DTYPE)
        return int(lib().orc_scene_create_tlas(self._h, _p(inst), inst.shape[0]))

    def delete_tlas(self, h: int):
        lib().orc_scene_delete_tlas(self._h, h)

    def bind_tlas(self, handles):
        h = _c(handles, np.uint32)
        lib().orc_scene_bind_tlas(self._h, _p(h), h.size)

    def build(self) -> int:
        return int(lib().orc_scene_build(self._h))

    def view(self) -> dict:
        v = _SceneView()
        lib().orc_scene_get_view(self._h, C.byref(v))
        dts = {"tlas_binding": np.dtype("u4"), "tlas_bvh_root": np.dtype("u4"), "tlas_bvh_forest": DEV_NODE_DTYPE,
               "tlas_data": DEV_INSTANCE_DTYPE, "tlas_bounding": TLAS_BOUNDING_DTYPE,
               "blas_meta_info": np.dtype(("u4", (2,))), "tri_bvh_root": GEOM_META_DTYPE, "tri_bvh_forest": DEV_NODE_DTYPE,
               "indices_redirect": np.dtype("u4"), "indices": np.dtype("u4"), "vertices": np.dtype(("f4", (3,)))}
        out = {}
        for name, dt in dts.items():
            n = int(getattr(v, "n_" + name)); ptr = getattr(v, name)
            if n == 0 or not ptr:
                out[name] = np.zeros(0, dt)
            else:
                buf = (C.c_char * (n * dt.itemsize)).from_address(ptr)
                out[name] = np.frombuffer(buf, dtype=dt).copy()
        out["balance_fallbacks"] = int(v.balance_fallbacks)
        out["balance_fallbacks_gt10"] = int(v.balance_fallbacks_gt10)
        return out

    def candidates(self, ray, ray_flags=0, cull_mask=0xFFFFFFFF, tlas_idx=0, cap=4096) -> np.ndarray:
        """brute force: every (instance, triangle) whose triangle test passes with the ray's original range (no BVH, no pruning)"""
        ray = _c(np.asarray(ray).reshape(1), RAY_DTYPE)
        out = np.zeros(cap, CANDIDATE_DTYPE)
        launch = Launch(ray_flags, cull_mask, tlas_idx, 0)
        L = lib()
        L.orc_scene_candidates.restype = C.c_uint64
        L.orc_scene_candidates.argtypes = [C.c_void_p, C.POINTER(Launch), C.c_void_p, C.c_void_p, C.c_uint64]
        n = L.orc_scene_candidates(self._h, C.byref(launch), _p(ray), _p(out), cap)
        return out[:min(int(n), cap)]

    def set_any_hit(self, programs=None, uniform_program=None, hit_group_any=None, sbt_ray_offset=0, sbt_ray_stride=1):
        """Any-hit shaders as data (oracle_scene.c "any-hit") for the traces that follow.  ``programs``: list of
        (kind, behavior, otherwise, mask, value, distance); ``uniform_program``: index used for all non-opaque geometry, or
        ``hit_group_any``: the any_hit handle of every SBT hit group (selection as in trace_task.rs:189-203).  No arguments: off."""
        L = lib()
        L.orc_scene_set_any_hit.restype = C.c_int
        L.orc_scene_set_any_hit.argtypes = [C.c_void_p, C.c_void_p]
        if programs is None:
            L.orc_scene_set_any_hit(self._h, None)
            return
        prog = np.zeros(len(programs), ANYHIT_PROGRAM_DTYPE)
        for k, p in enumerate(programs):
            prog[k] = tuple(p) + (0, 0)
        groups = np.zeros(0, np.uint32) if hit_group_any is None else _c(hit_group_any, np.uint32)
        setup = _AnyHitSetup(2 if hit_group_any is not None else 1, 0 if uniform_program is None else uniform_program, _p(prog), len(programs),
                             _p(groups) if groups.size else None, groups.size, sbt_ray_offset, sbt_ray_stride)
        L.orc_scene_set_any_hit(self._h, C.byref(setup))

    def trace_unpruned(self, rays, ray_flags=0, cull_mask=0xFFFFFFFF, tlas_idx=0, n_threads=1) -> np.ndarray:
        """NOT the reference: closest candidate of a walk that never shrinks its range (order-free model, oracle_scene.c)"""
        rays = _c(rays, RAY_DTYPE)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        launch = Launch(ray_flags, cull_mask, tlas_idx, 0)
        L = lib()
        L.orc_scene_trace_unpruned.restype = C.c_int
        L.orc_scene_trace_unpruned.argtypes = [C.c_void_p, C.POINTER(Launch), C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        if L.orc_scene_trace_unpruned(self._h, C.byref(launch), _p(rays), rays.shape[0], _p(hits), n_threads) != 0:
            raise RuntimeError("oracle trace failed (scene not built?)")
        return hits

    def trace_ordered_model(self, rays, ray_flags=0, cull_mask=0xFFFFFFFF, tlas_idx=0):
        """NOT the reference: scalar model of the CUDA kernel's ordered traversal + tie re-walk (oracle_scene.c); returns
        (hits, rays resolved by the clamped re-walk, rays resolved by the whole-range walk)"""
        rays = _c(rays, RAY_DTYPE)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        launch = Launch(ray_flags, cull_mask, tlas_idx, 0)
        stats = np.zeros(2, np.uint64)
        L = lib()
        L.orc_scene_trace_ordered_model.restype = C.c_int
        L.orc_scene_trace_ordered_model.argtypes = [C.c_void_p, C.POINTER(Launch), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        if L.orc_scene_trace_ordered_model(self._h, C.byref(launch), _p(rays), rays.shape[0], _p(hits), _p(stats)) != 0:
            raise RuntimeError("oracle trace failed (scene not built?)")
        return hits, int(stats[0]), int(stats[1])

    def trace(self, rays, ray_flags=0, cull_mask=0xFFFFFFFF, tlas_idx=0, n_threads=1, want_counters=True):
        rays = _c(rays, RAY_DTYPE)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        launch = Launch(ray_flags, cull_mask, tlas_idx, 0)
        ctr = Counters()
        rc = lib().orc_scene_trace(self._h, C.byref(launch), _p(rays), rays.shape[0], _p(hits), C.byref(ctr), n_threads)
        if rc != 0:
            raise RuntimeError(f"oracle trace failed rc={rc} (scene not built?)")
        return (hits, ctr.as_dict()) if want_counters else hits

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_scene_free(self._h)
            self._h = None


TOPOLOGY_POINT_LIST, TOPOLOGY_LINE_LIST, TOPOLOGY_LINE_STRIP, TOPOLOGY_TRIANGLE_LIST, TOPOLOGY_TRIANGLE_STRIP = range(5)


def _pick_args(positions, indices):
    pos = _c(positions, np.float32).reshape(-1, 3)
    idx = None if indices is None else _c(indices, np.uint32).reshape(-1)
    return pos, idx, (None if idx is None else _p(idx)), (0 if idx is None else idx.size)


def pick_nearest(positions, indices, topology, rays, tolerance=0.0, face_side=FACE_DOUBLE, n_threads=1) -> np.ndarray:
    """ray_intersect_nearest over every primitive (content/mesh/core/src/feature/intersection.rs:11-17), oracle_pick.c"""
    pos, idx, ip, ni = _pick_args(positions, indices)
    rays = _c(rays, RAY_DTYPE)
    out = np.zeros(rays.shape[0], MESH_HIT_DTYPE)
    L = lib()
    L.orc_pick_nearest.restype = None
    L.orc_pick_nearest.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
    L.orc_pick_nearest(_p(pos), pos.shape[0], ip, ni, topology, tolerance, face_side, _p(rays), rays.shape[0], _p(out), n_threads)
    return out


def pick_all(positions, indices, topology, ray, tolerance=0.0, face_side=FACE_DOUBLE) -> np.ndarray:
    """ray_intersect_all of one ray: every hit in primitive order (feature/intersection.rs:6-10)"""
    pos, idx, ip, ni = _pick_args(positions, indices)
    ray = _c(np.asarray(ray).reshape(1), RAY_DTYPE)
    L = lib()
    L.orc_pick_all.restype = C.c_uint64
    L.orc_pick_all.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
    n = int(L.orc_pick_all(_p(pos), pos.shape[0], ip, ni, topology, tolerance, face_side, _p(ray), None, 0))
    out = np.zeros(n, MESH_HIT_DTYPE)
    if n:
        L.orc_pick_all(_p(pos), pos.shape[0], ip, ni, topology, tolerance, face_side, _p(ray), _p(out), n)
    return out


def pick_primitive_count(n_positions, n_indices, has_indices, topology) -> int:
    L = lib()
    L.orc_pick_primitive_count.restype = C.c_uint64
    L.orc_pick_primitive_count.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    return int(L.orc_pick_primitive_count(n_positions, n_indices, 1 if has_indices else 0, topology))


def workgroup_inclusive_scan(x, workgroup: int) -> np.ndarray:
    x = _c(x, np.uint32); out = np.zeros_like(x)
    lib().orc_workgroup_inclusive_scan_u32(_p(x), x.size, workgroup, _p(out))
    return out


def inclusive_scan(x) -> np.ndarray:
    x = _c(x, np.uint32); out = np.zeros_like(x)
    lib().orc_inclusive_scan_u32(_p(x), x.size, _p(out))
    return out


def stream_compaction(x, keep):
    x = _c(x, np.uint32); keep = _c(keep, np.uint8); out = np.zeros_like(x)
    n = lib().orc_stream_compaction_u32(_p(x), _p(keep), x.size, _p(out))
    return out, int(n)


def shuffle_move(x, target, moved=None) -> np.ndarray:
    x = _c(x, np.uint32); target = _c(target, np.uint32)
    moved = np.ones(x.size, np.uint8) if moved is None else _c(moved, np.uint8)
    out = np.zeros_like(x)
    lib().orc_shuffle_move_u32(_p(x), _p(target), _p(moved), x.size, _p(out))
    return out


def mat4_compose(a, b) -> np.ndarray:
    a = _c(a, np.float32).reshape(16); b = _c(b, np.float32).reshape(16); out = np.zeros(16, np.float32)
    lib().orc_mat4_compose(_p(a), _p(b), _p(out))
    return out


def mat4_inverse_or_identity(m) -> np.ndarray:
    m = _c(m, np.float32).reshape(16); out = np.zeros(16, np.float32)
    lib().orc_mat4_inverse_or_identity(_p(m), _p(out))
    return out


def mat4_mul_vec4(m, v) -> np.ndarray:
    m = _c(m, np.float32).reshape(16); v = _c(v, np.float32).reshape(4); out = np.zeros(4, np.float32)
    lib().orc_mat4_mul_vec4(_p(m), _p(v), _p(out))
    return out
