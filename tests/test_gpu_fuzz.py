"""Randomised parity: seeded triangle soups (degenerate, duplicated and sliver triangles, indexed and not), several BLASes with
several geometries (some AABB geometry, some non-opaque), instances under random affine transforms (non-uniform and negative
scale, large translations, a singular matrix), several TLASes, rays with un-normalised, axis-aligned and zero-component
directions, origins on vertices and inside the scene, random ranges, flags and cull masks.  Every whole hit record must equal
the oracle's, for both kernels (NaN barycentrics of needle-triangle hits compare as NaN == NaN, helpers.canonical_nan).

These scenes are full of what the flattener classifies as irregular (needles, singular transforms, multi-geometry BLASes under
the reference's blas_box indexing): the reference's answer there depends on its visiting order, and the ordered kernel has to
recognise every ray that can reach such a thing and hand it to the reference-order walk."""
import numpy as np
import pytest

from rendiation_b200 import api, scenes as S

import helpers

pytestmark = pytest.mark.gpu
f32 = np.float32


def _soup(rng, n_tri, indexed, needles=None, scale=1.0, repair=False):
    """triangles inside the unit cube; a few duplicates (exact ties), zero-area and (needles: None = n/20) needle triangles"""
    centres = rng.uniform(-1, 1, (n_tri, 1, 3))
    tri = ((centres + rng.normal(0, rng.choice([0.02, 0.1, 0.4]), (n_tri, 3, 3))) * scale).astype(f32)
    k = max(1, n_tri // 20)
    tri[rng.integers(0, n_tri, k)] = tri[rng.integers(0, n_tri, k)]                 # duplicates
    d = rng.integers(0, n_tri, k); tri[d, 2] = tri[d, 1]                              # zero area
    if repair:  # no accidental near-needles (what the flattener would call irregular): give them a right angle
        e1, e2 = (tri[:, 1] - tri[:, 0]).astype(np.float64), (tri[:, 2] - tri[:, 0]).astype(np.float64)
        uu, vv, uv = (e1 * e1).sum(1), (e2 * e2).sum(1), (e1 * e2).sum(1)
        bad = (uu * vv - uv * uv < 1e-3 * uu * vv) & (uu > 0) & (vv > 0) & ((tri[:, 1] != tri[:, 2]).any(1))
        perp = np.stack([-e1[:, 1], e1[:, 0], np.zeros(n_tri)], 1)
        perp[(perp == 0).all(1)] = [1.0, 0.0, 0.0]
        tri[bad, 2] = (tri[bad, 0] + perp[bad]).astype(f32)
    s = rng.integers(0, n_tri, k if needles is None else needles)
    tri[s, 2] = tri[s, 0] + (tri[s, 1] - tri[s, 0]) * f32(0.5) + f32(1e-6)            # needles
    if not indexed:
        return np.ascontiguousarray(tri.reshape(-1, 3)), None
    pos = tri.reshape(-1, 3)
    idx = np.arange(pos.shape[0], dtype=np.uint32)
    # share some vertices between triangles
    share = rng.integers(0, idx.size, idx.size // 4)
    idx[share] = idx[rng.integers(0, idx.size, share.size)]
    return np.ascontiguousarray(pos), idx


def _transform(rng, kind):
    T, Sc, mul = S.mat4_translate, S.mat4_scale, S.mat4_mul
    rot = mul(mul(S.mat4_rotate_x(rng.uniform(-3, 3)), S.mat4_rotate_y(rng.uniform(-3, 3))), S.mat4_rotate_z(rng.uniform(-3, 3)))
    if kind == 0:
        sc = Sc(*rng.uniform(0.3, 3.0, 3))
    elif kind == 1:
        sc = Sc(*(rng.uniform(0.3, 3.0, 3) * rng.choice([-1.0, 1.0], 3)))  # mirrored: the winding flips
    elif kind == 2:
        sc = Sc(rng.uniform(0.5, 2), 0.0, rng.uniform(0.5, 2))               # singular -> inverse_or_identity
    else:
        sc = Sc(1, 1, 1)
    t = T(*rng.uniform(-6, 6, 3)) if kind != 3 else T(*(rng.uniform(-1, 1, 3) * 1000.0))
    return mul(mul(t, rot), sc)


def _scene(seed, devices=(0,), profile="hostile"):
    """hostile: everything at once — nearly every instance is irregular and most TLASes end up in the reference-order kernel;
    regular: soups with duplicates (exact ties) and zero-area triangles under well-conditioned (also mirrored) transforms, nothing
             irregular — every ray stays in the ordered kernel;
    mixed:   regular plus a BLAS with a few needles (leaf-level hand-over) and a few singular instances (instance-level)"""
    rng = np.random.default_rng(seed)
    sp = helpers.ScenePair(devices=devices)
    if profile != "hostile":
        return _tame_scene(sp, rng, profile)
    blases = []
    for _ in range(rng.integers(1, 4)):
        geoms = []
        for _ in range(rng.integers(1, 4)):
            if rng.random() < 0.15:
                geoms.append((rng.uniform(-1, 1, (4, 6)).astype(f32), None, 1, True))  # AABB geometry: accepted, ignored
                continue
            pos, idx = _soup(rng, int(rng.integers(1, 400)), indexed=rng.random() < 0.7)
            geoms.append((pos, idx, int(rng.choice([0, 1, 1, 3]))))                      # some non-opaque geometry
        blases.append(sp.blas(geoms))
    tl = []
    for _ in range(rng.integers(1, 4)):
        inst = []
        for i in range(rng.integers(1, 40)):
            kind = int(rng.choice([0, 0, 0, 1, 2, 3], p=[0.3, 0.2, 0.2, 0.2, 0.05, 0.05]))
            inst.append(S.make_instance(_transform(rng, kind), int(rng.choice(blases)), custom_index=int(rng.integers(0, 1 << 24)),
                                        mask=int(rng.choice([0xFFFFFFFF, 0x1, 0x2, 0xF0, 0])), flags=int(rng.choice([0, 0, 0, 1, 2, 4, 8, 3])),
                                        sbt_offset=int(rng.integers(0, 8))))
        tl.append(sp.tlas(np.concatenate(inst)))
    if rng.random() < 0.3 and len(blases) > 1:
        pass  # (deleting a referenced BLAS makes the reference panic; covered by test_errors_are_status_codes_not_aborts)
    sp.bind(tl)
    return sp.build(), len(tl), rng


def _tame_scene(sp, rng, profile):
    blases = []
    n_blas = int(rng.integers(2, 5))
    for b in range(n_blas):
        needles = 2 if (profile == "mixed" and b == 0) else 0
        pos, idx = _soup(rng, int(rng.integers(50, 600)), indexed=False, needles=needles, repair=True)
        geoms = [(pos, idx, int(rng.choice([0, 1, 1, 3])))]
        # The reference keeps one box per GEOMETRY but looks it up by BLAS handle (naive/mod.rs:239,273): only the last BLAS
        # may have more than one geometry without shifting the others' boxes, and it is boxed by its FIRST geometry alone
        if b == n_blas - 1:
            pos2, idx2 = _soup(rng, int(rng.integers(10, 100)), indexed=False, needles=0, scale=0.25, repair=True)
            inside = np.abs(pos2).max() < np.abs(pos).max(0).min() * 0.5
            if inside and (pos.min(0) < pos2.min(0)).all() and (pos.max(0) > pos2.max(0)).all():
                geoms.append((pos2, idx2, 1))
        blases.append(sp.blas(geoms))
    tl = []
    for _ in range(int(rng.integers(1, 3))):
        inst = []
        for i in range(int(rng.integers(12, 60))):
            kind = int(rng.choice([0, 1, 4]))
            tr = _transform(rng, kind if kind != 4 else 0)
            blas = int(rng.choice(blases))
            if profile == "mixed":  # at most IRREGULAR_LIST_MAX irregular instances per TLAS, or the whole TLAS is routed
                if i < 3:
                    tr = _transform(rng, 2)                   # singular
                blas = blases[0] if i in (3, 4) else int(rng.choice(blases[1:]))  # blases[0] holds the needles
            inst.append(S.make_instance(tr, blas, custom_index=int(rng.integers(0, 1 << 24)),
                                        mask=int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 0x1, 0xF0])), flags=int(rng.choice([0, 0, 0, 1, 2, 4, 8, 3])),
                                        sbt_offset=int(rng.integers(0, 8))))
        tl.append(sp.tlas(np.concatenate(inst)))
    sp.bind(tl)
    return sp.build(), len(tl), rng


def _rays(rng, n):
    o = rng.uniform(-8, 8, (n, 3)).astype(f32)
    target = rng.uniform(-6, 6, (n, 3)).astype(f32)
    d = (target - o).astype(f32)
    d *= rng.choice([1.0, 1.0, 0.01, 37.0], (n, 1)).astype(f32)          # not normalised: t is in units of |d|
    k = n // 10
    ax = rng.integers(0, n, k); d[ax] = 0; d[ax, rng.integers(0, 3, k)] = rng.choice([-1.0, 1.0], k)   # axis aligned (two zero components)
    z = rng.integers(0, n, k); d[z, rng.integers(0, 3, k)] = 0.0                                         # one zero component
    nz = rng.integers(0, n, k // 4); d[nz, rng.integers(0, 3, k // 4)] = -0.0                            # ... and a negative zero
    o[rng.integers(0, n, k)] = np.round(o[rng.integers(0, n, k)])                                        # origins on a lattice
    tmin = rng.choice([0.0, 0.0, 1e-3, 0.5], n).astype(f32)
    tmax = rng.choice([1e30, 100.0, 2.0, 0.75], n).astype(f32)
    rays = S.make_rays(o, d, 0.0, 1.0)
    rays["tmin"], rays["tmax"] = tmin, tmax
    # a sprinkling of rays no renderer should produce, which must still come out as the reference computes them: NaN / infinite
    # components, a zero direction, an empty or inverted range, a negative near
    w = rng.integers(0, n, 64)
    rays["ox"][w[0:6]] = np.nan; rays["dy"][w[6:12]] = np.nan; rays["oz"][w[12:16]] = np.inf; rays["dx"][w[16:20]] = -np.inf
    for f in ("dx", "dy", "dz"):
        rays[f][w[20:28]] = 0.0
    rays["tmin"][w[28:36]] = 5.0; rays["tmax"][w[28:36]] = 1.0
    rays["tmin"][w[36:44]] = 0.5; rays["tmax"][w[36:44]] = 0.5
    rays["tmin"][w[44:52]] = -3.0
    rays["tmax"][w[52:58]] = np.inf; rays["tmin"][w[58:64]] = np.nan
    return rays


@pytest.mark.parametrize("profile", ["regular", "mixed", "hostile"])
@pytest.mark.parametrize("seed", range(8))
def test_random_scenes_rays_and_flags(seed, profile):
    sp, n_tlas, rng = _scene(1000 + seed, profile=profile)
    stats = sp.p.build_stats()
    if profile == "regular":
        assert stats["irregular_triangles"] == 0 and stats["irregular_instances"] == 0, stats
    if profile == "mixed":
        assert stats["irregular_triangles"] > 0 and stats["irregular_instances"] >= 3, stats
    rays = _rays(rng, 12000)
    flag_sets = [0, api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, api.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES, api.RAY_FLAG_FORCE_OPAQUE,
                 api.RAY_FLAG_CULL_OPAQUE, api.RAY_FLAG_CULL_NON_OPAQUE | api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES,
                 api.RAY_FLAG_SKIP_TRIANGLES, api.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH,
                 api.RAY_FLAG_FORCE_NON_OPAQUE | api.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES]
    total_hits = 0
    for trial in range(5):
        flags = int(flag_sets[int(rng.integers(0, len(flag_sets)))])
        mask = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 0x3, 0xF0]))
        tlas_idx = int(rng.integers(0, n_tlas + 1))  # n_tlas = one past the binding: every ray misses
        want, wctr = sp.o.trace(rays, ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx, n_threads=4)
        got = sp.p.trace_closest_batch(rays, ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx)
        rep = helpers.compare_hits(got, want)
        assert rep["bit_identical"], (seed, trial, hex(flags), hex(mask), tlas_idx, rep)
        got_ref, gctr = sp.p.trace_counted(rays, ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx)
        assert helpers.identical_hits(got_ref, want) and gctr == wctr, (seed, trial, hex(flags), gctr, wctr)
        total_hits += rep["hits"]
    assert total_hits > 0
