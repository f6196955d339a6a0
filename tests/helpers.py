"""Shared scene builders for the parity tests: the same description is fed to the oracle and to the product."""
from __future__ import annotations

import numpy as np

import oracle
from rendiation_b200 import api, scenes as S

CULL_BACK = api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES


class ScenePair:
    """Builds the oracle scene and (optionally) the product scene from one description."""

    def __init__(self, devices=(0,), product=True):
        self.o = oracle.Scene()
        self.p = api.NaiveSahBVHSystem(devices=devices) if product else None
        self.blas_sources = []
        self.tlas_sources = []

    def blas(self, geometries):
        """geometries: list of (positions, indices|None, flags[, is_aabb])"""
        ho = self.o.create_blas(geometries)
        if self.p is not None:
            hp = self.p.create_bottom_level_acceleration_structure(
                [api.BottomLevelAccelerationStructureBuildSource(g[0], g[1], g[2], aabbs=(len(g) > 3 and g[3])) for g in geometries])
            assert hp.id == ho
        self.blas_sources.append(geometries)
        return ho

    def tlas(self, instances):
        self.tlas_sources.append(instances)
        ho = self.o.create_tlas(instances)
        if self.p is not None:
            hp = self.p.create_top_level_acceleration_structure(instances)
            assert hp.id == ho
        return ho

    def bind(self, handles):
        self.o.bind_tlas(handles)
        if self.p is not None:
            self.p.bind_tlas(handles)

    def build(self):
        rc = self.o.build()
        assert rc == 0, rc
        if self.p is not None:
            self.p.commit()
        return self


def single_mesh_scene(pos, idx, transform, devices=(0,), product=True, flags=1):
    sp = ScenePair(devices, product)
    b = sp.blas([(pos, idx.reshape(-1), flags)])
    t = sp.tlas(S.make_instance(transform, b))
    sp.bind([t])
    return sp.build()


def sphere_c1(devices=(0,), product=True, seg=64):
    pos, idx = S.uv_sphere_mesh(seg, seg)
    m = S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5))
    return single_mesh_scene(pos, idx, m, devices, product), (pos, idx, m)


def torus_scene(seg, devices=(0,), product=True):
    pos, idx = S.torus_mesh(seg, seg, 1.0, 0.35)
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
    return single_mesh_scene(pos, idx, m, devices, product), (pos, idx, m)


def reference_fixture(devices=(0,), product=True):
    """init_default_acceleration_structure (geometry/naive/test.rs:9-225): 3 BLAS, 5 TLAS; returns pair + tlas handles."""
    sp = ScenePair(devices, product)
    spos, sidx = S.uv_sphere_mesh(32, 16)
    tpos, tidx = S.torus_mesh(32, 8, 1.0, 0.2)
    sphere = sp.blas([(spos, sidx.reshape(-1), 1)])
    torus = sp.blas([(tpos, tidx.reshape(-1), 1)])
    cube = sp.blas([(S.CUBE_POSITION, S.CUBE_INDEX, 1)])
    T, Sc, Ry, Rz, Rx, mul = S.mat4_translate, S.mat4_scale, S.mat4_rotate_y, S.mat4_rotate_z, S.mat4_rotate_x, S.mat4_mul
    pi = np.float32(np.pi)
    src0 = [S.make_instance(T(i * 1.5, j * 1.5, -10.0), cube) for i in range(-2, 3) for j in range(-2, 3)]
    src0.append(S.make_instance(mul(T(0, 4.5, -10), Sc(5, 1, 1)), cube))
    src0.append(S.make_instance(mul(mul(T(0, -4.5, -10), Ry(pi)), Sc(5, 1, 1)), cube))
    src0.append(S.make_instance(mul(mul(T(4.5, -4.5, -10), Ry(pi * np.float32(0.5))), Sc(5, 1, 1)), cube))
    src0.append(S.make_instance(mul(mul(T(-4.5, -4.5, -10), Ry(pi * np.float32(-0.5))), Sc(5, 1, 1)), cube))
    tlas0 = sp.tlas(np.concatenate(src0))
    src1 = []
    for i in range(6):
        angle = np.float32(i) / np.float32(6.0) * pi * np.float32(2.0)
        s, c = np.float32(np.sin(angle)), np.float32(np.cos(angle))
        src1.append(S.make_instance(mul(mul(T(s * 4, c * 4, -5), Rz(-angle)), Sc(3, 0.5, 0.5)), cube))
    tlas1 = sp.tlas(np.concatenate(src1))
    tlas2 = sp.tlas(S.make_instance(mul(mul(T(0, 0, -10), Sc(5, 5, 5)), Rx(-0.5)), torus))
    src3 = [S.make_instance(T(i * 1.5, j * 1.5, -8.0 + k * 1.5), cube) for i in range(-2, 3) for j in range(-2, 3) for k in range(-2, 3)]
    tlas3 = sp.tlas(np.concatenate(src3))
    tlas4 = sp.tlas(S.make_instance(mul(T(0, 0, -10), Sc(5, 5, 5)), sphere))
    handles = [tlas0, tlas1, tlas2, tlas3, tlas4]
    sp.bind(handles)
    return sp.build(), handles


def canonical_nan(hits: np.ndarray) -> np.ndarray:
    """Hit records with every NaN replaced by one quiet NaN.  A needle triangle whose determinant cancels to zero yields a hit
    with finite t and NaN barycentrics (0 * inf) in the reference; which NaN is hardware-defined: x86 SSE generates 0xFFC00000
    and propagates operand payloads, the GPU always returns 0x7FFFFFFF.  Everything else stays bit for bit."""
    out = hits.copy()
    for f in out.dtype.names:
        if out.dtype[f] == np.float32:   # (t, u, v) of rdn_hit, (px, py, pz, distance) of rdn_mesh_hit
            bits = out[f].view(np.uint32)
            bits[np.isnan(out[f])] = 0x7FC00000
    return out


def identical_hits(got: np.ndarray, want: np.ndarray) -> bool:
    """whole 32-byte records equal, NaNs compared as NaN == NaN (see canonical_nan)"""
    return canonical_nan(got).tobytes() == canonical_nan(want).tobytes()


def compare_hits(got: np.ndarray, want: np.ndarray, rel_tol: float = 1e-6):
    """Parity report per the north star: ids bit-exact, t/u/v within rel_tol, near-ties reported separately."""
    assert got.shape == want.shape
    id_fields = ["primitive_id", "geometry_id", "instance_id", "instance_custom_id", "hit_kind"]
    ids_equal = np.ones(got.shape[0], bool)
    for f in id_fields:
        ids_equal &= got[f] == want[f]
    both_hit = (got["instance_id"] != 0xFFFFFFFF) & (want["instance_id"] != 0xFFFFFFFF)
    tw = want["t"].astype(np.float64)
    tg = got["t"].astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(tg - tw) / np.maximum(np.abs(tw), 1e-30)
    near_tie = (~ids_equal) & both_hit & (rel < rel_tol)
    hard_mismatch = (~ids_equal) & ~near_tie
    bits_equal = identical_hits(got, want)
    exact = ids_equal & both_hit
    max_rel_t = float(rel[exact].max()) if exact.any() else 0.0
    max_abs_uv = 0.0
    if exact.any():
        max_abs_uv = float(max(np.abs(got["u"][exact] - want["u"][exact]).max(), np.abs(got["v"][exact] - want["v"][exact]).max()))
    return {"n": int(got.shape[0]), "ids_exact": int(ids_equal.sum()), "near_ties": int(near_tie.sum()),
            "hard_mismatch": int(hard_mismatch.sum()), "bit_identical": bool(bits_equal), "max_rel_t": max_rel_t,
            "max_abs_uv": max_abs_uv, "hits": int((want["instance_id"] != 0xFFFFFFFF).sum())}


# ---------------------------------------------------------------------------------------------------------------
# numpy mirror of the ordered kernel's refill test (traverse.cu meets_irregular_instance), over the product's
# flattened arrays: which rays are handed to the reference-order walk because they can reach something irregular
IRREGULAR_ROUTE_ALL = 0xFFFFFFFF
IRREGULAR_WHOLE_BIT = 1 << 31


def _slab(o, inv, tmin, tmax, bmin, bmax):
    """intersect_ray_aabb_cpu, vectorised over rays (f32, fmin/fmax NaN semantics)"""
    with np.errstate(all="ignore"):
        t0 = ((bmin - o) * inv).astype(np.float32)
        t1 = ((bmax - o) * inv).astype(np.float32)
        near = np.fmax.reduce(np.fmin(t0, t1), axis=1)
        far = np.fmin.reduce(np.fmax(t0, t1), axis=1)
        return (near <= far) & (tmin < far) & (near < tmax)


def suspect_rays(arrays: dict, rays: np.ndarray, tlas_idx: int, cull_mask: int):
    """(route_all, mask): route_all = the bound TLAS takes the reference-order kernel as a whole; mask[i] = ray i meets an
    irregular instance (or an irregular leaf of its BLAS) with its original range."""
    n = rays.shape[0]
    binding = arrays["tlas_binding"]
    if tlas_idx >= binding.size or binding[tlas_idx] >= arrays["tlas_root"].shape[0]:
        return False, np.zeros(n, bool)
    _, _, start, count = (int(x) for x in arrays["tlas_root"][binding[tlas_idx]][:4])
    if count == IRREGULAR_ROUTE_ALL:
        return True, np.ones(n, bool)
    f32 = np.float32
    o = np.stack([rays["ox"], rays["oy"], rays["oz"]], 1).astype(f32)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], 1).astype(f32)
    tmin, tmax = rays["tmin"].astype(f32), rays["tmax"].astype(f32)
    with np.errstate(all="ignore"):
        inv = (f32(1.0) / d).astype(f32)
    mask = np.zeros(n, bool)
    for entry in arrays["irregular_instances"][start:start + count]:
        slot = int(entry) & ~IRREGULAR_WHOLE_BIT
        tb = arrays["tlas_bounding"][slot]
        meets = _slab(o, inv, tmin, tmax, tb["world_min"], tb["world_max"]) & bool(cull_mask & int(tb["mask"]))
        if int(entry) & IRREGULAR_WHOLE_BIT:
            mask |= meets
            continue
        rec = arrays["instances"][slot]
        blas = int(rec["blas"])
        if blas >= arrays["blas_meta"].shape[0]:
            continue
        l0, ln = int(arrays["blas_meta"][blas][2]), int(arrays["blas_meta"][blas][3])
        m = rec["transform_inv"].reshape(4, 4)  # m[c] = column c
        with np.errstate(all="ignore"):
            p = [((((o[:, 0] * m[0][i]).astype(f32) + (o[:, 1] * m[1][i]).astype(f32)).astype(f32) + (o[:, 2] * m[2][i]).astype(f32)).astype(f32)
                  + f32(f32(1.0) * m[3][i])).astype(f32) for i in range(4)]
            bo = np.stack([(p[i] / p[3]).astype(f32) for i in range(3)], 1)
            d0 = np.stack([(((d[:, 0] * m[0][i]).astype(f32) + (d[:, 1] * m[1][i]).astype(f32)).astype(f32) + (d[:, 2] * m[2][i]).astype(f32)).astype(f32)
                           for i in range(3)], 1)
            len2 = (((d0[:, 0] * d0[:, 0]).astype(f32) + (d0[:, 1] * d0[:, 1]).astype(f32)).astype(f32) + (d0[:, 2] * d0[:, 2]).astype(f32)).astype(f32)
            scaling = np.sqrt(len2).astype(f32)
            bd = np.where((len2 > 0)[:, None], (d0 * (f32(1.0) / scaling)[:, None]).astype(f32), d0)
            inv_bd = (f32(1.0) / bd).astype(f32)
        leaf = np.zeros(n, bool)
        with np.errstate(all="ignore"):
            smin, smax = (tmin * scaling).astype(f32), (tmax * scaling).astype(f32)
        for lb in arrays["irregular_leaf_boxes"][l0:l0 + ln]:
            leaf |= _slab(bo, inv_bd, smin, smax, lb["bmin"], lb["bmax"])
        mask |= meets & leaf
    return False, mask
