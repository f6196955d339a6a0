"""The flattener's regularity classification (accel.cpp) on the CPU: which triangles / instances make the reference's answer
depend on its visiting order, and that the rays the ordered kernel would keep for itself are exactly rays whose reference
result is order-free.  Uses host-only scenes (no GPU): the classification lives in the flattened arrays."""
import numpy as np
import pytest

from rendiation_b200 import api, scenes as S

import helpers
import test_gpu_fuzz as fuzz

f32 = np.float32
ROUTE_ALL = helpers.IRREGULAR_ROUTE_ALL
WHOLE = helpers.IRREGULAR_WHOLE_BIT


def _host(build):
    sp = build(devices=(), product=True)
    return sp[0] if isinstance(sp, tuple) else sp


def test_benchmark_scenes_are_regular():
    for build in (helpers.sphere_c1, lambda **k: helpers.torus_scene(96, **k), helpers.reference_fixture):
        sp = _host(build)
        st = sp.p.build_stats()
        assert st["irregular_triangles"] == 0 and st["irregular_instances"] == 0 and st["reference_routed_tlas"] == 0, st
        a = sp.p.arrays()
        assert not a["tlas_root"][:, 3].any() and not a["blas_meta"][:, 3].any()
        assert a["irregular_instances"].size == 0 and a["irregular_leaf_boxes"].size == 0


def test_high_tessellation_sphere_stays_regular():
    # thin (not needle) triangles next to the poles and exactly degenerate pole triangles: both harmless
    pos, idx = S.uv_sphere_mesh(708, 708)
    sp = helpers.single_mesh_scene(pos, idx, S.mat4_translate(0, 0, -10), devices=(), product=True)
    assert sp.p.build_stats()["irregular_triangles"] == 0


def _needle():
    a = np.array([0.0, 0.0, 0.0], f32)
    b = np.array([1.0, 0.0, 0.0], f32)
    return np.stack([a, b, (a + (b - a) * f32(0.5) + f32(1e-6)).astype(f32)])


def test_needle_triangle_marks_its_leaf_only():
    pos, idx = S.torus_mesh(24, 12, 1.0, 0.3)
    pos = np.concatenate([pos, _needle()]).astype(f32)
    idx = np.concatenate([idx.reshape(-1), np.arange(3, dtype=np.uint32) + (pos.shape[0] - 3)])
    sp = helpers.ScenePair(devices=(), product=True)
    b = sp.blas([(pos, idx, 1)])
    far = sp.blas([(S.CUBE_POSITION, S.CUBE_INDEX, 1)])
    inst = [S.make_instance(S.mat4_translate(0, 0, -5), b)] + [S.make_instance(S.mat4_translate(4.0 * i, 6, -5), far) for i in range(-3, 4)]
    sp.bind([sp.tlas(np.concatenate(inst))])
    sp.build()
    st = sp.p.build_stats()
    assert st["irregular_triangles"] == 1 and st["irregular_instances"] == 1 and st["reference_routed_tlas"] == 0, st
    a = sp.p.arrays()
    assert tuple(a["blas_meta"][b][2:]) == (0, 1) and tuple(a["blas_meta"][far][2:]) == (1, 0)
    assert a["irregular_leaf_boxes"].shape[0] == 1
    start, count = (int(x) for x in a["tlas_root"][0][2:4])
    assert count == 1 and not (int(a["irregular_instances"][start]) & WHOLE)
    assert int(a["instances"][int(a["irregular_instances"][start])]["blas"]) == b
    # only rays through the needle's leaf box are kept for the reference-order walk
    rays = S.pinhole_rays(64, 64, tmin=0.0, tmax=100.0)
    route_all, mask = helpers.suspect_rays(a, rays, 0, 0xFFFFFFFF)
    assert not route_all and 0 < mask.sum() < 0.1 * rays.size


def _cube_scene(instances_of):
    sp = helpers.ScenePair(devices=(), product=True)
    b = sp.blas([(S.CUBE_POSITION, S.CUBE_INDEX, 1)])
    sp.bind([sp.tlas(np.concatenate(instances_of(b)))])
    return sp.build()


def test_irregular_instances_are_listed_or_the_tlas_is_routed():
    T, Sc, mul = S.mat4_translate, S.mat4_scale, S.mat4_mul
    # one singular transform among regular ones: listed as a whole
    sp = _cube_scene(lambda b: [S.make_instance(T(3.0 * i, 0, -10), b) for i in range(-4, 5)] + [S.make_instance(mul(T(0, 5, -10), Sc(1, 0, 1)), b)])
    a = sp.p.arrays()
    start, count = (int(x) for x in a["tlas_root"][0][2:4])
    assert count == 1 and int(a["irregular_instances"][start]) & WHOLE
    assert sp.p.build_stats()["irregular_instances"] == 1
    # the only instance is singular / more than IRREGULAR_LIST_MAX irregular instances: whole TLAS in reference order
    for make in (lambda b: [S.make_instance(Sc(1, 0, 1), b)],
                 lambda b: [S.make_instance(mul(T(3.0 * i, 0, -10), Sc(0, 1, 1)), b) for i in range(9)] + [S.make_instance(T(3.0 * i, 4, -10), b) for i in range(40)]):
        sp = _cube_scene(make)
        assert int(sp.p.arrays()["tlas_root"][0][3]) == ROUTE_ALL and sp.p.build_stats()["reference_routed_tlas"] == 1
    # projective and non-finite matrices
    for bad in (np.array([1, 0, 0, 0.1, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, -10, 1], f32),
                np.array([1, 0, 0, 0, 0, np.inf, 0, 0, 0, 0, 1, 0, 0, 0, -10, 1], f32)):
        sp = _cube_scene(lambda b: [S.make_instance(bad, b)])
        assert int(sp.p.arrays()["tlas_root"][0][3]) == ROUTE_ALL


def test_reference_blas_box_indexing_makes_instances_irregular():
    """naive/mod.rs:239,273: blas_box holds one entry per GEOMETRY but is indexed by BLAS handle — after a two-geometry BLAS
    every later BLAS gets the box of the wrong geometry, so hits of its instances need not lie inside their world boxes"""
    cube = (S.CUBE_POSITION, S.CUBE_INDEX, 1)
    small = ((S.CUBE_POSITION * f32(0.1)).astype(f32), S.CUBE_INDEX, 1)
    sp = helpers.ScenePair(devices=(), product=True)
    two = sp.blas([cube, small])          # blas_box = [cube, small]
    other = sp.blas([cube])               # handle 1 -> blas_box[1] = the SMALL cube's box
    sp.bind([sp.tlas(np.concatenate([S.make_instance(S.mat4_translate(3.0 * i, 0, -10), two) for i in range(10)] +
                                    [S.make_instance(S.mat4_translate(0, 4, -10), other)]))])
    sp.build()
    a = sp.p.arrays()
    start, count = (int(x) for x in a["tlas_root"][0][2:4])
    listed = [int(e) & ~WHOLE for e in a["irregular_instances"][start:start + count]]
    assert count == 1 and int(a["instances"][listed[0]]["blas"]) == other


@pytest.mark.parametrize("profile", ["regular", "mixed", "hostile"])
@pytest.mark.parametrize("seed", range(8))
def test_rays_kept_by_the_ordered_kernel_have_order_free_results(seed, profile):
    """On the hostile fuzz scenes: wherever the reference's pre-order walk and an order-free walk of the same trees disagree
    (beyond an exact tie) the ray must be one the ordered kernel hands over (suspect), or the whole TLAS is routed."""
    sp, n_tlas, rng = fuzz._scene(1000 + seed, devices=(), profile=profile)
    rays = fuzz._rays(rng, 6000)
    arrays = sp.p.arrays()
    stats = sp.p.build_stats()
    if profile == "regular":
        assert stats["irregular_triangles"] == 0 and stats["irregular_instances"] == 0 and stats["reference_routed_tlas"] == 0, stats
    kept_total = 0
    for tlas_idx in range(n_tlas):
        for flags, mask in ((0, 0xFFFFFFFF), (api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, 0x3),
                            (api.RAY_FLAG_FORCE_OPAQUE | api.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES, 0xF0)):
            ref = sp.o.trace(rays, ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx, n_threads=4, want_counters=False)
            free = sp.o.trace_unpruned(rays, ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx, n_threads=4)
            route_all, suspect = helpers.suspect_rays(arrays, rays, tlas_idx, mask)
            differs = (ref["t"].view(np.uint32) != free["t"].view(np.uint32)) | ((ref["instance_id"] == 0xFFFFFFFF) != (free["instance_id"] == 0xFFFFFFFF))
            kept = ~suspect
            assert not (differs & kept).any(), (seed, tlas_idx, hex(flags), int((differs & kept).sum()), np.nonzero(differs & kept)[0][:5])
            kept_total += int(kept.sum())
    if profile != "hostile":
        assert kept_total > 0


def _slivers(rng, n, lo, hi, length=0.4):
    """triangles whose edges at v0 enclose an angle with sin^2 log-uniform in [lo, hi]"""
    c = rng.uniform(-1, 1, (n, 3))
    e1 = rng.normal(size=(n, 3)); e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    perp = np.cross(e1, rng.normal(size=(n, 3))); perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    theta = np.arcsin(np.sqrt(np.exp(rng.uniform(np.log(lo), np.log(hi), n))))
    e2 = (np.cos(theta)[:, None] * e1 + np.sin(theta)[:, None] * perp) * rng.uniform(0.5, 1.0, (n, 1))
    return np.stack([c, c + length * e1, c + length * e2], 1).astype(f32).reshape(-1, 3)


@pytest.mark.parametrize("band", [(2e-5, 1e-4), (1e-4, 1e-2), (1e-12, 1e-9)])
def test_needle_threshold_has_margin(band):
    """Where does the reference's answer start to depend on its visiting order?  Thousands of slivers per band, 150 K rays through them:
    none above the flattener's threshold (sin^2 = 1e-5; in a sweep the first order-dependent rays appear below 1e-6), plenty far
    below it — and there every triangle is flagged, so every ray is handed over."""
    rng = np.random.default_rng(11)
    pos = _slivers(rng, 12000, *band)
    sp = helpers.ScenePair(devices=(), product=True)
    b = sp.blas([(pos, None, 1)])
    sp.bind([sp.tlas(S.make_instance(S.mat4_identity(), b))])
    sp.build()
    n = 150000
    o = rng.uniform(-1.2, 1.2, (n, 3)).astype(f32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = S.make_rays(o, d.astype(f32), 0.0, 100.0)
    ref = sp.o.trace(rays, n_threads=8, want_counters=False)
    free = sp.o.trace_unpruned(rays, n_threads=8)
    differs = (ref["t"].view(np.uint32) != free["t"].view(np.uint32)) | ((ref["instance_id"] == 0xFFFFFFFF) != (free["instance_id"] == 0xFFFFFFFF))
    route_all, suspect = helpers.suspect_rays(sp.p.arrays(), rays, 0, 0xFFFFFFFF)
    st = sp.p.build_stats()
    assert int((ref["instance_id"] != 0xFFFFFFFF).sum()) > 5000
    if band[0] >= 2e-5:
        assert st["irregular_triangles"] == 0 and not route_all and not suspect.any()
        assert not differs.any(), int(differs.sum())
    else:
        assert st["irregular_triangles"] == 12000 and route_all
        assert int(differs.sum()) > 100


@pytest.mark.parametrize("profile", ["regular", "mixed", "hostile"])
@pytest.mark.parametrize("seed", range(8))
def test_ordered_traversal_algorithm_ends_on_the_reference_record(seed, profile):
    """A scalar model of the kernel's ordered traversal (near child first, pruning against the inflated bound, near-tie detection,
    clamped re-walk, whole-range fallback; oracle_scene.c orc_scene_trace_ordered_model) must produce the reference's whole hit record
    on every ray the kernel keeps for itself — the exactness argument of DESIGN.md §3, fuzzed on the CPU at a volume GPU time does
    not allow.  (Rays the flattener's classification hands over are walked in reference order by construction.)"""
    sp, n_tlas, rng = fuzz._scene(2000 + seed, devices=(), profile=profile)
    rays = fuzz._rays(rng, 20000)
    arrays = sp.p.arrays()
    checked = 0
    for tlas_idx in range(n_tlas):
        for flags, mask in ((0, 0xFFFFFFFF), (api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, 0xFFFFFFFF), (api.RAY_FLAG_FORCE_NON_OPAQUE | api.RAY_FLAG_CULL_FRONT_FACING_TRIANGLES, 0xF3)):
            route_all, suspect = helpers.suspect_rays(arrays, rays, tlas_idx, mask)
            if route_all:
                continue
            kept = ~suspect
            ref = sp.o.trace(rays[kept], ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx, n_threads=4, want_counters=False)
            got, ties, whole = sp.o.trace_ordered_model(rays[kept], ray_flags=flags, cull_mask=mask, tlas_idx=tlas_idx)
            assert helpers.identical_hits(got, ref), (seed, profile, tlas_idx, hex(flags), int((helpers.canonical_nan(got) != helpers.canonical_nan(ref)).sum()))
            checked += int(kept.sum())
    if profile != "hostile":
        assert checked > 0


def test_tie_window_on_stacks_of_near_duplicate_triangles():
    """five layers of every triangle, displaced along its normal by 5e-7 .. 5e-3 (straddling TIE_EPS = 1e-5 relative): a fifth of
    the hitting rays are near-ties; the modelled ordered traversal must still end on the reference's record for every ray
    (offline: 3.6 M rays, 740 K near-ties, 0 mismatches)"""
    rng = np.random.default_rng(21)
    n = 400
    c = rng.uniform(-1, 1, (n, 1, 3))
    tri = c + rng.normal(0, 0.25, (n, 3, 3))
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    layers = [tri]
    for _ in range(4):
        eps = np.exp(rng.uniform(np.log(1e-7), np.log(1e-3), (n, 1, 1))) * 5.0
        layers.append(tri + nrm[:, None, :] * eps * rng.choice([-1, 1], (n, 1, 1)))
    pos = np.concatenate(layers).astype(f32).reshape(-1, 3)
    sp = helpers.ScenePair(devices=(), product=True)
    b = sp.blas([(pos, None, 1)])
    T, Sc, Ry, mul = S.mat4_translate, S.mat4_scale, S.mat4_rotate_y, S.mat4_mul
    sp.bind([sp.tlas(np.concatenate([S.make_instance(mul(T(*rng.uniform(-2, 2, 3)), mul(Ry(rng.uniform(0, 3)), Sc(*rng.uniform(0.5, 2, 3)))), b)
                                     for _ in range(6)]))])
    sp.build()
    assert sp.p.build_stats()["irregular_instances"] == 0
    m = 40000
    o = rng.uniform(-5, 5, (m, 3)).astype(f32)
    d = rng.uniform(-2, 2, (m, 3)) - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = S.make_rays(o, d.astype(f32), 0.0, 100.0)
    for flags in (0, api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES):
        ref = sp.o.trace(rays, ray_flags=flags, n_threads=4, want_counters=False)
        got, ties, whole = sp.o.trace_ordered_model(rays, ray_flags=flags)
        assert helpers.identical_hits(got, ref)
        assert ties > 2000 and whole == 0
