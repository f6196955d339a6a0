"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed goldens.

Bar (BASELINE.json north_star): hit primitive/instance ids bit-exact, t/u/v within 1e-6 relative, near-ties counted
separately.  Because the kernels use the reference's arithmetic un-fused and resolve near-ties in reference order, the
expectation here is stronger: whole hit records bit-identical.
"""
import json
import os

import numpy as np
import pytest

import oracle
from rendiation_b200 import api, scenes as S

import helpers

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CTR = ("bvh_visit", "bvh_hit", "tri_visit", "tri_hit", "inst_visit", "ref_abort")
REL_TOL = 1e-6  # north star tolerance for t / barycentrics


def _report(name, rep):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity.jsonl", "a") as f:
        f.write(json.dumps({"case": name, **rep}) + "\n")


def _assert_parity(name, got, want, require_bits=True):
    rep = helpers.compare_hits(got, want, REL_TOL)
    _report(name, rep)
    assert rep["hard_mismatch"] == 0, rep
    assert rep["max_rel_t"] <= REL_TOL and rep["max_abs_uv"] <= REL_TOL, rep
    if require_bits:
        assert rep["near_ties"] == 0 and rep["bit_identical"], rep
    return rep


def _device_trace(sysm, rays, mode, **kw):
    import torch
    r = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
    h = torch.zeros((rays.shape[0], 32), dtype=torch.uint8, device="cuda")
    stats = sysm.trace_closest_device(r.data_ptr(), rays.shape[0], h.data_ptr(), stream=torch.cuda.current_stream().cuda_stream,
                                      mode=mode, want_stats=True, **kw)
    torch.cuda.synchronize()
    return h.cpu().numpy().view(api.HIT_DTYPE).reshape(-1), stats


# ---- goldens (no oracle needed at run time) -----------------------------------------------------------
@pytest.mark.parametrize("k", range(5))
def test_reference_fixture_goldens(k):
    g = np.load(os.path.join(GOLD, "reference_fixture_64.npz"))
    sp, _ = helpers.reference_fixture()
    rays = g["rays"]
    for name, flags in (("cull_back", 0x10), ("none", 0x00), ("first_hit", 0x14)):
        want = g[f"hits_tlas{k}_{name}"]
        got = sp.p.trace_closest_batch(rays, ray_flags=flags, tlas_idx=k)
        _assert_parity(f"fixture_tlas{k}_{name}_auto", got, want)
        got_ref, ctr = sp.p.trace_counted(rays, ray_flags=flags, tlas_idx=k)
        _assert_parity(f"fixture_tlas{k}_{name}_reforder", got_ref, want)
        assert [ctr[c] for c in CTR] == g[f"ctr_tlas{k}_{name}"].tolist(), (name, ctr)


def test_reference_test_output_file_from_the_gpu():
    """the file and the counter printout of the reference's test_cpu_triangle (tests/golden/trace_cpu.pbm, made by the oracle; a
    maintainer can byte-compare it with the Rust test's) come out of the CUDA path unchanged: ids from the ordered kernel, the four
    visit counters from the reference-order kernel"""
    import json
    import sys
    sys.path.insert(0, GOLD)
    import make_golden
    sp, _ = helpers.reference_fixture()

    def trace(rays, flags, tlas_idx):
        hits = sp.p.trace_closest_batch(rays, ray_flags=flags, tlas_idx=tlas_idx, grid_width=256)
        counted, ctr = sp.p.trace_counted(rays, ray_flags=flags, tlas_idx=tlas_idx)
        assert counted.tobytes() == hits.tobytes()
        return hits, ctr

    text, counters = make_golden.trace_cpu_pbm(None, trace)
    assert text == open(os.path.join(GOLD, "trace_cpu.pbm")).read()
    assert counters == json.load(open(os.path.join(GOLD, "trace_cpu_counters.json")))


def test_c1_golden_path_b_and_a():
    g = np.load(os.path.join(GOLD, "c1_sphere_96.npz"))
    sp, (pos, idx, m) = helpers.sphere_c1()
    got = sp.p.trace_closest_batch(g["rays"], ray_flags=0x10, grid_width=96)
    _assert_parity("c1_golden_b", got, g["hits_b"])
    wpos = S.mat4_apply_point(m, pos)
    bvh = api.build_bvh_for_abstract_mesh(wpos, idx, api.SAH(4), api.TreeBuildOption(50, 2))
    got_a = api.intersect_nearest_bvh(wpos, idx, g["rays"], bvh, api.FACE_DOUBLE)
    assert got_a.tobytes() == g["hits_a"].tobytes()


# ---- oracle comparisons on seeded inputs -------------------------------------------------------------
@pytest.mark.parametrize("mode", [api.TRACE_AUTO, api.TRACE_REFERENCE_ORDER])
def test_c1_sphere_full_size(mode):
    """BASELINE config 1: 1024x1024 coherent primary rays vs the 64x64-segment sphere"""
    sp, _ = helpers.sphere_c1()
    rays = S.pinhole_rays(1024, 1024, 0.0, 100.0)
    want, ctr = sp.o.trace(rays, ray_flags=0x10, n_threads=os.cpu_count() or 4)
    assert ctr["ref_abort"] == 0
    got, stats = _device_trace(sp.p, rays, mode, ray_flags=0x10, grid_width=1024)
    rep = _assert_parity(f"c1_full_mode{mode}", got, want)
    assert rep["hits"] > 200000
    if mode == api.TRACE_REFERENCE_ORDER:
        _, gctr = sp.p.trace_counted(rays[:65536], ray_flags=0x10)
        _, octr = sp.o.trace(rays[:65536], ray_flags=0x10, n_threads=4)
        assert gctr == octr


def test_torus_medium_primary_and_bounce():
    """configs 2/3 at a size the oracle finishes in seconds: 256x256-segment torus (131k tris), 640x360 primary rays +
    one incoherent cosine-weighted bounce per hit (no culling)"""
    sp, (pos, idx, m) = helpers.torus_scene(256)
    rays = S.pinhole_rays(640, 360, 0.01, 100.0, aspect_correct=True)
    nt = os.cpu_count() or 4
    want, ctr = sp.o.trace(rays, ray_flags=0x10, n_threads=nt)
    got = sp.p.trace_closest_batch(rays, ray_flags=0x10, grid_width=640)
    rep = _assert_parity("torus256_primary", got, want)
    assert rep["hits"] > 20000 and ctr["ref_abort"] == 0
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], -1)
    hit = want["instance_id"] != 0xFFFFFFFF
    normals = np.zeros((rays.shape[0], 3), np.float32)
    normals[hit] = S.geometric_normals(pos, idx, want["primitive_id"][hit], m, d[hit])
    brays, _ = S.bounce_rays(rays, want, normals)
    bwant, bctr = sp.o.trace(brays, ray_flags=0, n_threads=nt)
    bgot = sp.p.trace_closest_batch(brays, ray_flags=0)
    brep = _assert_parity("torus256_bounce", bgot, bwant)
    assert brep["hits"] > 1000 and bctr["ref_abort"] == 0


def test_instanced_scene_ids_exact():
    """config 4 shape at reduced size: 20x20 transform-instanced copies of a 64x64 sphere, instance ids exact"""
    pos, idx = S.uv_sphere_mesh(64, 64)
    sp = helpers.ScenePair()
    b = sp.blas([(pos, idx.reshape(-1), 1)])
    t = sp.tlas(S.instance_grid(20, 20, b, spacing=3.5, z=-60.0))
    sp.bind([t])
    sp.build()
    rays = S.pinhole_rays(512, 512, 0.0, 1000.0)
    want, ctr = sp.o.trace(rays, ray_flags=0x10, n_threads=os.cpu_count() or 4)
    got = sp.p.trace_closest_batch(rays, ray_flags=0x10, grid_width=512)
    rep = _assert_parity("instanced_20x20", got, want)
    assert rep["hits"] > 10000 and len(np.unique(want["instance_custom_id"])) > 50


_FULL_SIZE = {}   # scene and oracle results of the full-size test, shared by its variants


@pytest.mark.parametrize("variant", [0, 100, 110, 120, 130, 140])
def test_c2_c3_full_size(variant, monkeypatch):
    """BASELINE configs 2 and 3 at full size: 1920x1080 primary rays vs the 1,002,528-triangle torus, then one incoherent
    cosine-weighted bounce ray per hit; whole hit records bit-identical to the oracle on every ray — with the shipped choice of
    loops (0: tile history on the grid, top-up on the two-million-ray list, late work sharing on the bounce list) and with every
    loop forced on every launch (100 plain, 110 / 120 work sharing, 130 top-up, 140 top-up then sharing)"""
    monkeypatch.setenv("RDN_ORDERED_VARIANT", str(variant))
    if not _FULL_SIZE:
        sp, (pos, idx, m) = helpers.torus_scene(708)
        rays = S.pinhole_rays(1920, 1080, 0.01, 100.0, aspect_correct=True)
        _FULL_SIZE.update(sp=sp, pos=pos, idx=idx, m=m, rays=rays, primary=sp.o.trace(rays, ray_flags=0x10, n_threads=os.cpu_count() or 4))
    sp, pos, idx, m, rays = (_FULL_SIZE[k] for k in ("sp", "pos", "idx", "m", "rays"))
    assert idx.size // 3 == 1002528
    st = sp.p.build_stats()
    assert st["irregular_triangles"] == 0 and st["irregular_instances"] == 0 and st["balance_fallbacks_gt10"] == 0
    nt = os.cpu_count() or 4
    want, ctr = _FULL_SIZE["primary"]
    got, stats = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10, grid_width=1920)
    rep = _assert_parity("c2_full", got, want)
    assert rep["hits"] > 300000 and ctr["ref_abort"] == 0 and stats["whole_range_rewalks"] == 0
    # the same frame as a LIST of more than two million rays: the launch takes the loop that tops thinned-out warps up with new rays
    lgot, _ = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10)
    _assert_parity("c2_full_as_list", lgot, want)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], -1)
    hit = want["instance_id"] != 0xFFFFFFFF
    normals = np.zeros((rays.shape[0], 3), np.float32)
    normals[hit] = S.geometric_normals(pos, idx, want["primitive_id"][hit], m, d[hit])
    brays, _ = S.bounce_rays(rays, want, normals)
    if "bounce" not in _FULL_SIZE:
        _FULL_SIZE["bounce"] = sp.o.trace(brays, ray_flags=0, n_threads=nt)
    bwant, bctr = _FULL_SIZE["bounce"]
    bgot, bstats = _device_trace(sp.p, brays, api.TRACE_AUTO, ray_flags=0)
    brep = _assert_parity("c3_full", bgot, bwant)
    assert brep["n"] == int(hit.sum()) and brep["hits"] > 10000 and bctr["ref_abort"] == 0 and bstats["whole_range_rewalks"] == 0


def test_c4_instanced_full_size():
    """BASELINE config 4 at full size: 10,000 transform-instanced copies of a 100,352-triangle sphere, 1920x1080 primary rays;
    (instance_id, instance_custom_id, primitive_id) and everything else bit-identical, near-ties included"""
    pos, idx = S.uv_sphere_mesh(224, 224)
    assert idx.size // 3 == 100352
    sp = helpers.ScenePair()
    b = sp.blas([(pos, idx.reshape(-1), 1)])
    sp.bind([sp.tlas(S.instance_grid(100, 100, b, 3.5, -200.0))])
    sp.build()
    st = sp.p.build_stats()
    assert st["irregular_triangles"] == 0 and st["irregular_instances"] == 0 and st["reference_routed_tlas"] == 0
    rays = S.pinhole_rays(1920, 1080, 0.0, 1000.0, aspect_correct=True)
    want, ctr = sp.o.trace(rays, ray_flags=0x10, n_threads=os.cpu_count() or 4)
    got, stats = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10, grid_width=1920)
    rep = _assert_parity("c4_full", got, want)
    assert rep["hits"] > 400000 and len(np.unique(want["instance_custom_id"])) > 5000
    assert stats["tie_rays"] > 0 and stats["whole_range_rewalks"] == 0  # touching spheres of neighbouring instances: real near-ties


def test_flags_masks_multi_geometry_and_edge_cases():
    cube = (S.CUBE_POSITION, S.CUBE_INDEX, 1)
    non_opaque = (S.CUBE_POSITION + np.array([2.0, 0, 0], np.float32), S.CUBE_INDEX, 0)
    nonindexed = (S.CUBE_POSITION[S.CUBE_INDEX] + np.array([-2.0, 0, 0], np.float32), None, 1)
    aabb = (np.array([[0, 0, 0, 1, 1, 1]], np.float32), None, 1, True)
    sp = helpers.ScenePair()
    b = sp.blas([cube, non_opaque, nonindexed, aabb])
    empty = sp.blas([(np.zeros((0, 3), np.float32), np.zeros(0, np.uint32), 1)])
    T, Sc, mul = S.mat4_translate, S.mat4_scale, S.mat4_mul
    inst = np.concatenate([
        S.make_instance(T(0, 0, -6), b, custom_index=7, mask=0x1),
        S.make_instance(mul(T(0, 2.5, -6), Sc(-1, 1, 1)), b, custom_index=8, mask=0x2),           # det < 0 -> FLIP_FACING
        S.make_instance(T(0, -2.5, -6), b, custom_index=9, mask=0x4, flags=api.GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE),
        S.make_instance(T(0, 0, -9), b, custom_index=10, mask=0x8, flags=api.GEOMETRY_INSTANCE_FORCE_NO_OPAQUE),
        S.make_instance(T(0, 0, -3), empty, custom_index=11, mask=0xFF),
    ])
    t_main = sp.tlas(inst)
    t_deleted = sp.tlas(S.make_instance(T(0, 0, -4), b))
    sp.o.delete_tlas(t_deleted)
    sp.p.delete_top_level_acceleration_structure(api.TlasHandle(t_deleted))
    sp.bind([t_main, t_deleted])
    sp.build()
    rays = S.pinhole_rays(160, 160, 0.0, 100.0)
    for flags in (0x00, 0x10, 0x20, 0x40, 0x80, 0x01 | 0x40, 0x02 | 0x80, 0x100, 0x04, 0x14):
        for mask in (0xFFFFFFFF, 0x1, 0x6, 0x0):
            want, _ = sp.o.trace(rays, ray_flags=flags, cull_mask=mask, tlas_idx=0, n_threads=4)
            got = sp.p.trace_closest_batch(rays, ray_flags=flags, cull_mask=mask, tlas_idx=0)
            _assert_parity(f"flags{flags:#x}_mask{mask:#x}", got, want)
    # a deleted TLAS and an out-of-range tlas_idx both miss everywhere
    for tl in (1, 5):
        got = sp.p.trace_closest_batch(rays[:256], tlas_idx=tl)
        assert (got["instance_id"] == 0xFFFFFFFF).all() and (got["t"] == 100.0).all()
    # empty and ragged launches
    assert sp.p.trace_closest_batch(rays[:0]).shape == (0,)
    for n in (1, 31, 33, 1000):
        want, _ = sp.o.trace(rays[:n], ray_flags=0x10)
        assert sp.p.trace_closest_batch(rays[:n], ray_flags=0x10).tobytes() == want.tobytes()
    # rays starting inside geometry / tight ranges
    tight = rays.copy()
    tight["tmin"], tight["tmax"] = 5.4, 5.6
    want, _ = sp.o.trace(tight, ray_flags=0)
    _assert_parity("tight_range", sp.p.trace_closest_batch(tight, ray_flags=0), want)


def test_exact_ties_resolve_like_the_reference():
    """coplanar duplicated geometry: every hit is an exact tie; the reference keeps the LAST accepted in its visit order"""
    pos, idx = S.uv_sphere_mesh(24, 24)
    sp = helpers.ScenePair()
    b1 = sp.blas([(pos, idx.reshape(-1), 1)])
    b2 = sp.blas([(pos, idx.reshape(-1), 1), (pos, idx.reshape(-1), 1)])  # duplicated geometry inside one BLAS
    m = S.mat4_mul(S.mat4_translate(0, 0, -6), S.mat4_scale(2, 2, 2))
    t = sp.tlas(np.concatenate([S.make_instance(m, b1, custom_index=1), S.make_instance(m, b1, custom_index=2),
                                S.make_instance(m, b2, custom_index=3)]))
    sp.bind([t])
    sp.build()
    rays = S.pinhole_rays(256, 256, 0.0, 100.0)
    want, ctr = sp.o.trace(rays, ray_flags=0x10, n_threads=4)
    got, stats = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10, grid_width=256)
    rep = _assert_parity("exact_ties", got, want)
    assert stats["tie_rays"] >= rep["hits"] > 1000  # every hit went through the exact tie resolution


def test_degenerate_triangles_are_ignored():
    """zero-area pole triangles without culling: NaN t is never reported (the reference's CPU path aborts there)"""
    pos, idx = S.uv_sphere_mesh(16, 16)
    m = S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5))
    sp = helpers.single_mesh_scene(pos, idx, m)
    rays = S.pinhole_rays(128, 128, 0.0, 100.0)
    want, ctr = sp.o.trace(rays, ray_flags=0)
    got = sp.p.trace_closest_batch(rays, ray_flags=0)
    assert not np.isnan(got["t"]).any()
    _assert_parity("degenerate_no_cull", got, want)


def test_path_a_space_query_matches_oracle():
    pos, idx = S.torus_mesh(96, 64)
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
    wpos = S.mat4_apply_point(m, pos)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([wpos[tri].min(1), wpos[tri].max(1)], 1)
    rays = S.pinhole_rays(200, 200, 0.0, 100.0)
    for strat_o, strat_p, opt in ((oracle.STRATEGY_SAH, api.SAH(4), (50, 2)), (oracle.STRATEGY_SAH, api.SAH(4), (10, 50)),
                                  (oracle.STRATEGY_BALANCE, api.BalanceTree(), (15, 10))):
        ob = oracle.FlattenBVH(boxes, strat_o, 4, *opt)
        pb = api.build_bvh_for_abstract_mesh(wpos, idx, strat_p, api.TreeBuildOption(*opt))
        for side in (api.FACE_FRONT, api.FACE_BACK, api.FACE_DOUBLE):
            want = ob.query_nearest(wpos, idx, rays, side, 4)
            got = api.intersect_nearest_bvh(wpos, idx, rays, pb, side)
            assert got.tobytes() == want.tobytes(), (opt, side)
            assert want["hit"].sum() > 1000


def test_path_a_device_resident_query_and_edge_cases():
    """upload once / query device buffers; n = 0; a query before upload is an error code; empty tree answers none"""
    import torch
    pos, idx = S.uv_sphere_mesh(24, 24)  # has zero-area pole triangles: DdN == 0 -> none, as in the reference
    wpos = S.mat4_apply_point(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), pos)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([wpos[tri].min(1), wpos[tri].max(1)], 1)
    rays = S.pinhole_rays(160, 120, 0.0, 100.0)
    ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, 50, 2)
    pb = api.build_bvh_for_abstract_mesh(wpos, idx, api.SAH(4), api.TreeBuildOption(50, 2))
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
    d_out = torch.full((rays.shape[0], 32), 0xAB, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    with pytest.raises(api.RdnError):
        api.intersect_nearest_bvh_device(pb, d_rays.data_ptr(), rays.shape[0], d_out.data_ptr(), api.FACE_DOUBLE, st)
    api.upload_bvh(pb, wpos, idx, 0)
    api.intersect_nearest_bvh_device(pb, d_rays.data_ptr(), 0, d_out.data_ptr(), api.FACE_DOUBLE, st)  # nothing to do
    for side in (api.FACE_FRONT, api.FACE_BACK, api.FACE_DOUBLE):
        api.intersect_nearest_bvh_device(pb, d_rays.data_ptr(), rays.shape[0], d_out.data_ptr(), side, st)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(api.MESH_HIT_DTYPE).reshape(-1)
        want = ob.query_nearest(wpos, idx, rays, side, 4)
        assert got.tobytes() == want.tobytes(), side
    # brute force over the same mesh agrees (the BVH only prunes)
    brute = oracle.brute_query_nearest(wpos, idx, rays, api.FACE_DOUBLE, 4)
    assert np.array_equal(brute["hit"], want["hit"]) and np.array_equal(brute["distance"][want["hit"] == 1], want["distance"][want["hit"] == 1])
    # empty mesh -> empty tree -> every ray is OptionalNearest::none()
    eb = api.build_bvh_for_abstract_mesh(wpos[:3], np.zeros(0, np.uint32), api.SAH(4), api.TreeBuildOption(50, 2))
    api.upload_bvh(eb, wpos[:3], np.zeros(0, np.uint32), 0)
    api.intersect_nearest_bvh_device(eb, d_rays.data_ptr(), rays.shape[0], d_out.data_ptr(), api.FACE_DOUBLE, st)
    torch.cuda.synchronize()
    assert not d_out.any().item()


def test_path_a_list_query_matches_oracle():
    """intersect_list_bvh: ragged CSR result, per-ray order = the reference's visiting order; sizes via the two-call protocol"""
    pos, idx = S.torus_mesh(64, 48)
    wpos = S.mat4_apply_point(S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5)), pos)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([wpos[tri].min(1), wpos[tri].max(1)], 1)
    rays = S.pinhole_rays(150, 130, 0.0, 100.0)
    for opt in ((50, 2), (10, 50)):
        ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, *opt)
        pb = api.build_bvh_for_abstract_mesh(wpos, idx, api.SAH(4), api.TreeBuildOption(*opt))
        for side in (api.FACE_FRONT, api.FACE_DOUBLE):
            woff, whits = ob.query_list(wpos, idx, rays, side)
            goff, ghits = api.intersect_list_bvh(wpos, idx, rays, pb, side)
            assert np.array_equal(goff, woff) and ghits.tobytes() == whits.tobytes(), (opt, side)
            assert whits.size > 2000
    goff, ghits = api.intersect_list_bvh(wpos, idx, rays[:0], pb)  # no rays
    assert goff.tolist() == [0] and ghits.size == 0
    away = S.pinhole_rays(8, 8, 0.0, 100.0, origin=(0.0, 0.0, -40.0))  # looking away from the mesh: empty lists
    goff, ghits = api.intersect_list_bvh(wpos, idx, away, pb)
    assert not goff.any() and ghits.size == 0


# ---- wavefront queue compaction ---------------------------------------------------------------------
def test_compaction_known_answers_and_random():
    s = api.NaiveSahBVHSystem()
    x = np.array([1, 0, 1, 0, 1, 1, 0], np.uint32)  # stream_compaction.rs:100-125
    out, n = s.compact_u32(x, x == 1)
    assert out.tolist() == [1, 1, 1, 1, 0, 0, 0] and n == 4
    rng = np.random.default_rng(3)
    for size in (0, 1, 7, 2047, 2048, 2049, 70, 100003, 5_000_000):
        for p in (0.0, 0.3, 1.0):
            vals = rng.integers(0, 2 ** 32, size, dtype=np.uint32)
            keep = (rng.random(size) < p).astype(np.uint8) * rng.integers(1, 255, size, dtype=np.uint8)
            want, wn = oracle.stream_compaction(vals, keep)
            got, gn = s.compact_u32(vals, keep)
            assert gn == wn == int((keep != 0).sum()) and np.array_equal(got, want), (size, p)


def test_compaction_feeds_a_bounce_wave():
    """wavefront step: trace primaries, compact the indices of rays that hit (device side), gather those rays, trace again"""
    import torch
    sp, (pos, idx, m) = helpers.torus_scene(128)
    rays = S.pinhole_rays(256, 256, 0.01, 100.0)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
    d_hits = torch.zeros_like(d_rays)
    n = rays.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    sp.p.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=0x10, stream=st)
    inst = d_hits.view(torch.int32)[:, 5]
    keep = (inst != -1).to(torch.uint8)
    ids = torch.arange(n, dtype=torch.int32, device="cuda")
    out = torch.empty_like(ids)
    out_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    sp.p.compact_u32_device(ids.data_ptr(), keep.data_ptr(), n, out.data_ptr(), out_n.data_ptr(), stream=st)
    torch.cuda.synchronize()
    want_hits, _ = sp.o.trace(rays, ray_flags=0x10, n_threads=4)
    want_ids = np.nonzero(want_hits["instance_id"] != 0xFFFFFFFF)[0]
    k = int(out_n.item())
    assert k == want_ids.size and np.array_equal(out[:k].cpu().numpy(), want_ids) and (out[k:] == 0).all()


# ---- replication ---------------------------------------------------------------------------------------
def test_blob_adoption_gives_identical_results():
    import torch
    sp, _ = helpers.reference_fixture()
    ptr, nbytes = sp.p.blob()
    # copy the blob like an NCCL broadcast would deliver it, then adopt it in a second scene object
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    from cuda.bindings import runtime as cudart
    (err,) = cudart.cudaMemcpy(src.data_ptr(), ptr, nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert int(err) == 0
    other = api.NaiveSahBVHSystem()
    other.adopt_blob(src.data_ptr(), nbytes)
    del src
    rays = S.pinhole_rays(96, 96, 0.0, 100.0)
    for k in range(5):
        a = sp.p.trace_closest_batch(rays, ray_flags=0x10, tlas_idx=k)
        b = other.trace_closest_batch(rays, ray_flags=0x10, tlas_idx=k)
        assert a.tobytes() == b.tobytes()


def test_in_process_multi_device_scene_shards_host_rays():
    """rdn_rt_scene_create(n_devices = 2): the blob is replicated by peer copy on commit and the host-buffer trace deals ray chunks
    round-robin to the devices; the assembled result is the single-device result (= the oracle's), and each device answers
    device-resident traces from its own replica."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    pos, idx = S.torus_mesh(128, 96, 1.0, 0.35)
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
    sp = helpers.single_mesh_scene(pos, idx, m, devices=(0, 1))
    rays = S.pinhole_rays(1024, 640, 0.01, 100.0)  # 655,360 rays: three 256 Ki chunks -> both devices work
    want = sp.o.trace(rays, ray_flags=helpers.CULL_BACK, n_threads=8, want_counters=False)
    got = sp.p.trace_closest_batch(rays, ray_flags=helpers.CULL_BACK, grid_width=1024)
    _assert_parity("two_devices_host_path", got, want)
    for di in (0, 1):
        with torch.cuda.device(di):
            r = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).to(f"cuda:{di}")
            h = torch.zeros_like(r)
            sp.p.trace_closest_device(r.data_ptr(), rays.shape[0], h.data_ptr(), ray_flags=helpers.CULL_BACK, grid_width=1024,
                                      stream=torch.cuda.current_stream(di).cuda_stream, device_index=di)
            torch.cuda.synchronize(di)
            assert h.cpu().numpy().view(api.HIT_DTYPE).reshape(-1).tobytes() == want.tobytes(), di


def test_back_to_back_launches_overlap_without_losing_rays():
    """Consecutive ordered launches on one stream overlap their tails (programmatic dependent launch, two alternating scratch
    sets guarded by an epoch gate).  Every launch writes its own pre-filled buffer; all must equal a synchronised launch, which
    equals the oracle.  Also alternates two streams (the library orders them through an event) and mixes in a small launch
    (partial grid: no overlap requested) and a reference-order launch."""
    import torch
    sp, _ = helpers.torus_scene(160)
    W, H = 1280, 800  # 1,024,000 rays: the grid fills the GPU, so overlap is requested
    rays = S.pinhole_rays(W, H, 0.01, 100.0)
    n = rays.shape[0]
    want = sp.o.trace(rays, ray_flags=helpers.CULL_BACK, n_threads=8, want_counters=False)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
    ref = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    s0 = torch.cuda.Stream(); s1 = torch.cuda.Stream()
    sp.p.trace_closest_device(d_rays.data_ptr(), n, ref.data_ptr(), ray_flags=helpers.CULL_BACK, grid_width=W, stream=s0.cuda_stream)
    torch.cuda.synchronize()
    assert ref.cpu().numpy().view(api.HIT_DTYPE).reshape(-1).tobytes() == want.tobytes()

    def burst(streams, count, small_every=0, reforder_every=0, same_buffer=False):
        outs = [torch.full((n, 32), 0xAB, dtype=torch.uint8, device="cuda") for _ in range(count)]
        torch.cuda.synchronize()
        if same_buffer:  # every launch writes the SAME hit buffer: the library must refuse the overlap (stragglers of launch k
            outs = [outs[0]] * count  # would overwrite records of launch k + 1) and the result is still right
        for k, o in enumerate(outs):
            st = streams[k % len(streams)]
            if small_every and k % small_every == small_every - 1:
                m = 4096 * W // W  # a launch far too small to fill the GPU, in the middle of the burst
                sp.p.trace_closest_device(d_rays.data_ptr(), m, o.data_ptr(), ray_flags=helpers.CULL_BACK, stream=st.cuda_stream, overlap_previous=True)
                sp.p.trace_closest_device(d_rays.data_ptr() + m * 32, n - m, o.data_ptr() + m * 32, ray_flags=helpers.CULL_BACK, stream=st.cuda_stream,
                                          overlap_previous=True)
            elif reforder_every and k % reforder_every == reforder_every - 1:
                sp.p.trace_closest_device(d_rays.data_ptr(), n, o.data_ptr(), ray_flags=helpers.CULL_BACK, stream=st.cuda_stream,
                                          mode=api.TRACE_REFERENCE_ORDER, overlap_previous=True)
            else:
                sp.p.trace_closest_device(d_rays.data_ptr(), n, o.data_ptr(), ray_flags=helpers.CULL_BACK, grid_width=W, stream=st.cuda_stream,
                                          overlap_previous=True)
        torch.cuda.synchronize()
        assert sp.p.poll_errors(stream=streams[0].cuda_stream) == 0
        return [bool(torch.equal(o, ref)) for o in outs]

    assert all(burst([s0], 24))                                   # one stream: every launch overlaps its predecessor's tail
    assert all(burst([s0, s1], 16))                               # two streams alternating: ordered through the library's event
    assert all(burst([s0], 18, small_every=5, reforder_every=7))  # partial grids and the other kernel in between
    assert all(burst([s0], 12, same_buffer=True))                 # aliased hit buffers: overlap refused by the library


@pytest.mark.gpu
def test_tile_history_reorders_the_launch_not_the_results():
    """Device-resident grid launches take their tiles in the order built from the previous launch over a grid of the same size (long
    passes first).  Whatever the order — first launch (grid order), repeated frames, a moved camera over the same grid, another grid
    size in between, a grid whose tile count is no multiple of anything, history switched off — every launch gives the oracle's
    records, and every ray is traced exactly once (the hit buffers start out as 0xAB)."""
    import torch
    sp, _ = helpers.torus_scene(160)
    nt = os.cpu_count() or 4
    st = torch.cuda.current_stream().cuda_stream

    def frame(W, H, tmin, shift):
        rays = S.pinhole_rays(W, H, tmin, 100.0, aspect_correct=True)
        rays["ox"] += np.float32(shift)
        return rays
    cases = [(640, 408, 0.0), (640, 408, 0.0), (640, 408, 0.35), (641, 403, 0.0), (640, 408, -0.2), (1021, 517, 0.0), (1021, 517, 0.1), (640, 408, 0.0)]
    for k, (W, H, shift) in enumerate(cases):
        rays = frame(W, H, 0.01, shift)
        want = sp.o.trace(rays, ray_flags=helpers.CULL_BACK, n_threads=nt, want_counters=False)
        r = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32).copy()).cuda()
        h = torch.full((rays.shape[0], 32), 0xAB, dtype=torch.uint8, device="cuda")
        sp.p.trace_closest_device(r.data_ptr(), rays.shape[0], h.data_ptr(), ray_flags=helpers.CULL_BACK, grid_width=W, stream=st)
        torch.cuda.synchronize()
        _assert_parity(f"tile_history_launch{k}", h.cpu().numpy().view(api.HIT_DTYPE).reshape(-1), want)


@pytest.mark.gpu
def test_pageable_buffers_go_through_the_staging_pipeline():
    """Ordinary caller memory (what the reference's API hands over) is staged chunk by chunk through page-locked buffers by several
    host threads: several chunks, buffers at odd addresses (the streaming copy aligns its stores itself), a page-locked array on
    one side only — always the oracle's records."""
    sp, _ = helpers.torus_scene(128)
    W, H = 1024, 640   # 655,360 rays: three chunks of the default size
    rays = S.pinhole_rays(W, H, 0.01, 100.0, aspect_correct=True)
    n = rays.shape[0]
    want = sp.o.trace(rays, ray_flags=helpers.CULL_BACK, n_threads=os.cpu_count() or 4, want_counters=False)
    for ray_shift, hit_shift in ((0, 0), (4, 12), (20, 8)):
        raw_rays = np.zeros(n * 32 + 64, np.uint8)
        raw_rays[ray_shift:ray_shift + n * 32] = rays.view(np.uint8).reshape(-1)
        raw_hits = np.full(n * 32 + 64, 0xAB, np.uint8)
        sp.p.trace_closest_host_ptr(raw_rays.ctypes.data + ray_shift, n, raw_hits.ctypes.data + hit_shift, ray_flags=helpers.CULL_BACK, grid_width=W)
        assert raw_hits[hit_shift:hit_shift + n * 32].tobytes() == want.tobytes(), (ray_shift, hit_shift)
        assert (raw_hits[:hit_shift] == 0xAB).all() and (raw_hits[hit_shift + n * 32:] == 0xAB).all()  # nothing written outside
    hits = np.zeros(n, api.HIT_DTYPE)
    with api.HostRegistration(hits):   # rays staged, hits copied straight into the page-locked array
        sp.p.trace_closest_host_ptr(rays.ctypes.data, n, hits.ctypes.data, ray_flags=helpers.CULL_BACK, grid_width=W)
    assert hits.tobytes() == want.tobytes()
    hits[:] = 0
    with api.HostRegistration(rays):   # ... and the other way round
        sp.p.trace_closest_host_ptr(rays.ctypes.data, n, hits.ctypes.data, ray_flags=helpers.CULL_BACK, grid_width=W)
    assert hits.tobytes() == want.tobytes()


@pytest.mark.gpu
def test_host_buffers_page_locked_by_their_owner():
    """rdn_rt_host_register / rdn_rt_host_alloc: the host-buffer trace on memory the caller page-locked (a registered numpy array,
    a library-allocated buffer) returns the records of the plain pageable call; registrations end with their owner."""
    import ctypes as C
    sp, _ = helpers.torus_scene(96)
    rays = S.pinhole_rays(512, 384, 0.01, 100.0)
    want = sp.p.trace_closest_batch(rays, ray_flags=helpers.CULL_BACK, grid_width=512)
    hits = np.zeros(rays.shape[0], api.HIT_DTYPE)
    with api.HostRegistration(rays), api.HostRegistration(hits):
        sp.p.trace_closest_host_ptr(rays.ctypes.data, rays.shape[0], hits.ctypes.data, ray_flags=helpers.CULL_BACK, grid_width=512)
    assert hits.tobytes() == want.tobytes()
    L = api.lib()
    p_rays, p_hits = C.c_void_p(), C.c_void_p()
    assert L.rdn_rt_host_alloc(rays.nbytes, C.byref(p_rays)) == 0 and L.rdn_rt_host_alloc(hits.nbytes, C.byref(p_hits)) == 0
    C.memmove(p_rays, rays.ctypes.data, rays.nbytes)
    sp.p.trace_closest_host_ptr(p_rays.value, rays.shape[0], p_hits.value, ray_flags=helpers.CULL_BACK, grid_width=512)
    assert C.string_at(p_hits, hits.nbytes) == want.tobytes()
    L.rdn_rt_host_free(p_rays); L.rdn_rt_host_free(p_hits)
    # the arrays are ordinary memory again: a fresh allocation may land on the same addresses and must still copy
    del rays, hits
    again = sp.p.trace_closest_batch(S.pinhole_rays(512, 384, 0.01, 100.0), ray_flags=helpers.CULL_BACK, grid_width=512)
    assert again.tobytes() == want.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 2, 9, 30, 40, 60, 61, 100, 110, 120, 130, 140])
def test_every_instantiation_of_the_ordered_kernel_is_bit_identical(variant, monkeypatch):
    """Every instantiation RDN_ORDERED_VARIANT can select — the shipped ones (0: grids plain, ray lists sharing work between lanes;
    100 / 110 force either for both launch kinds) and the experiments kept for A/B runs (2 the round-1 kernel, 9 the separate tie
    kernel, 30 128-bit loads, 40 the top of the trees staged in shared memory by TMA, 60 / 61 the four-box nodes) — gives the
    oracle's records on a grid launch, on the incoherent bounce list off its hits, and on an instanced scene with several
    instances per TLAS leaf.  (The variant is read at every launch; the flattener reads it for the four-box view.)"""
    if variant == 40 and os.environ.get("RDN_SIMT_EMU") == "1":
        pytest.skip("the TMA staging has no CPU stand-in")
    monkeypatch.setenv("RDN_ORDERED_VARIANT", str(variant))
    nt = os.cpu_count() or 4
    sp, (pos, idx, m) = helpers.torus_scene(192)
    W, H = 640, 400
    rays = S.pinhole_rays(W, H, 0.01, 100.0, aspect_correct=True)
    want = sp.o.trace(rays, ray_flags=0x10, n_threads=nt, want_counters=False)
    got, stats = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10, grid_width=W)
    _assert_parity(f"variant{variant}_grid", got, want)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], -1)
    hit = want["instance_id"] != 0xFFFFFFFF
    normals = np.zeros((rays.shape[0], 3), np.float32)
    normals[hit] = S.geometric_normals(pos, idx, want["primitive_id"][hit], m, d[hit])
    brays, _ = S.bounce_rays(rays, want, normals)
    bwant = sp.o.trace(brays, ray_flags=0, n_threads=nt, want_counters=False)
    bgot, _ = _device_trace(sp.p, brays, api.TRACE_AUTO, ray_flags=0)
    _assert_parity(f"variant{variant}_list", bgot, bwant)
    # instanced: 20 x 20 spheres, up to ten instances per TLAS leaf in the reference's tree
    spos, sidx = S.uv_sphere_mesh(48, 48)
    si = helpers.ScenePair((0,), True)
    b = si.blas([(spos, sidx.reshape(-1), 1)])
    si.bind([si.tlas(S.instance_grid(20, 20, b, 3.5, -60.0))])
    si.build()
    irays = S.pinhole_rays(480, 320, 0.0, 1000.0, aspect_correct=True)
    iwant = si.o.trace(irays, ray_flags=0x10, n_threads=nt, want_counters=False)
    igot, _ = _device_trace(si.p, irays, api.TRACE_AUTO, ray_flags=0x10, grid_width=480)
    rep = _assert_parity(f"variant{variant}_instanced", igot, iwant)
    assert rep["hits"] > 10000
    ilist, _ = _device_trace(si.p, irays, api.TRACE_AUTO, ray_flags=0x10)  # the same rays as a list (no grid hint)
    _assert_parity(f"variant{variant}_instanced_list", ilist, iwant)


def _non_opaque_scene():
    """a non-opaque torus (geometry flags 0) in front of an opaque sphere, two instances with different SBT record offsets"""
    tpos, tidx = S.torus_mesh(96, 48, 1.0, 0.35)
    spos, sidx = S.uv_sphere_mesh(48, 48)
    sp = helpers.ScenePair((0,), True)
    torus = sp.blas([(tpos, tidx.reshape(-1), 0)])                      # non-opaque: candidates go through the any-hit stage
    both = sp.blas([(spos, sidx.reshape(-1), 1), (tpos, tidx.reshape(-1), 0)])   # geometry 0 opaque, geometry 1 not
    T, Sc, Rx, mul = S.mat4_translate, S.mat4_scale, S.mat4_rotate_x, S.mat4_mul
    inst = np.concatenate([S.make_instance(mul(mul(T(-2.5, 0, -9), Sc(3, 3, 3)), Rx(-0.6)), torus, custom_index=11, sbt_offset=0),
                           S.make_instance(mul(T(3.0, 0, -11), Sc(2.2, 2.2, 2.2)), both, custom_index=22, sbt_offset=2)])
    sp.bind([sp.tlas(inst)])
    sp.build()
    return sp


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ignore_odd_primitives", "see_through_near", "accept_and_end_search", "end_search_without_accept", "from_sbt"])
def test_any_hit_stage_decides_over_candidates_of_non_opaque_geometry(case):
    """traverse_cpu.rs:164-192: candidates of non-opaque geometry are put to the any-hit stage; ACCEPT_HIT commits, END_SEARCH stops the
    traversal.  Stateless programs without END_SEARCH run in the ordered kernel, the others in reference order; whole records equal
    the oracle's either way, for grid launches, ray lists and the host-buffer call."""
    sp = _non_opaque_scene()
    A, E = api.ANYHIT_BEHAVIOR_ACCEPT_HIT, api.ANYHIT_BEHAVIOR_END_SEARCH
    programs = [(api.ANYHIT_PRIMITIVE_MASK, A, 0, 1, 0, 0.0),      # 0: odd primitives are holes
                (api.ANYHIT_MIN_DISTANCE, A, 0, 0, 0, 8.5),        # 1: nothing nearer than 8.5 is seen
                (api.ANYHIT_CONSTANT, A | E, 0, 0, 0, 0.0),        # 2: first candidate wins (= ACCEPT_FIRST_HIT_AND_END_SEARCH)
                (api.ANYHIT_PRIMITIVE_MASK, A, E, 3, 0, 0.0)]      # 3: every fourth primitive is a surface, any other one ends the search unseen
    sp.p.set_any_hit_programs(programs)
    kw = {}
    if case == "from_sbt":
        sbt = sp.p.create_sbt(2, 4, 1)
        sbt.config_hit_group(0, 0, 0, api.HitGroupShaderRecord(closest_hit=0, any_hit=0))   # instance 0 (record offset 0), geometry 0: program 0
        sbt.config_hit_group(1, 2, 0, api.HitGroupShaderRecord(closest_hit=0, any_hit=1))   # instance 1 (record offset 2), geometry 1 -> group 0 + 1 * 1 + 2 = 3: program 1
        sp.p.bind_sbt(sbt)
        groups = np.full(2 * 4 * 1, 0xFFFFFFFF, np.uint32)
        groups[0], groups[3] = 0, 1
        sp.o.set_any_hit(programs, hit_group_any=groups, sbt_ray_offset=0, sbt_ray_stride=1)
        kw = dict(any_hit=api.ANYHIT_FROM_SBT, sbt_ray=(0, 1))
    else:
        k = ["ignore_odd_primitives", "see_through_near", "accept_and_end_search", "end_search_without_accept"].index(case)
        sp.o.set_any_hit(programs, uniform_program=k)
        kw = dict(any_hit=k + 1)
    W, H = 320, 200
    rays = S.pinhole_rays(W, H, 0.0, 100.0, aspect_correct=True)
    staged = sp.o.trace(rays, ray_flags=0, n_threads=4, want_counters=False)
    for flags in (0x00, 0x10):
        want, wctr = sp.o.trace(rays, ray_flags=flags, n_threads=4)
        got, _ = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=flags, grid_width=W, **kw)
        _assert_parity(f"anyhit_{case}_{flags:#x}_grid", got, want)
        got, _ = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=flags, **kw)
        _assert_parity(f"anyhit_{case}_{flags:#x}_list", got, want)
        got = sp.p.trace_closest_batch(rays, ray_flags=flags, grid_width=W, **kw)
        _assert_parity(f"anyhit_{case}_{flags:#x}_host", got, want)
        got, gctr = sp.p.trace_counted(rays, ray_flags=flags, **kw)
        assert got.tobytes() == want.tobytes() and gctr == wctr
    # the stage changes answers (otherwise the test proves nothing), and FORCE_OPAQUE switches it off
    forced, _ = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x01, grid_width=W, **kw)
    _assert_parity(f"anyhit_{case}_force_opaque", forced, sp.o.trace(rays, ray_flags=0x01, n_threads=4, want_counters=False))
    sp.o.set_any_hit()
    base = sp.o.trace(rays, ray_flags=0, n_threads=4, want_counters=False)
    assert base.tobytes() != staged.tobytes()
    assert forced.tobytes() == sp.o.trace(rays, ray_flags=0x01, n_threads=4, want_counters=False).tobytes()  # (FORCE_OPAQUE: no stage)


@pytest.mark.gpu
def test_tlas_only_update_patches_the_resident_scene_in_place():
    """rdn_rt_tlas_update: moving instances (same TLAS handle, new transforms) rebuilds the TLAS part alone and patches the device blob
    in place; the traced records equal a scene built from scratch with the new transforms (the oracle), the BLAS arrays are not
    touched, and a later BLAS mutation goes back to a full build."""
    import time
    spos, sidx = S.uv_sphere_mesh(48, 48)
    sp = helpers.ScenePair((0,), True)
    b = sp.blas([(spos, sidx.reshape(-1), 1)])
    inst0 = S.instance_grid(30, 30, b, 3.5, -80.0)
    t = sp.tlas(inst0)
    sp.bind([t]); sp.build()
    rays = S.pinhole_rays(480, 320, 0.0, 1000.0, aspect_correct=True)
    before, _ = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10, grid_width=480)
    _assert_parity("tlas_update_before", before, sp.o.trace(rays, ray_flags=0x10, n_threads=4, want_counters=False))
    tri_before = sp.p.array(8).tobytes()   # RDN_ARRAY_TRIANGLES
    assert sp.p.build_stats()["tlas_only_commits"] == 0
    # every instance moves: a different grid spacing and depth, same count
    inst1 = S.instance_grid(30, 30, b, 4.25, -95.0)
    t0 = time.perf_counter()
    sp.p.update_top_level_acceleration_structure(api.TlasHandle(t), inst1)
    sp.p.commit()
    dt_ms = (time.perf_counter() - t0) * 1e3
    st = sp.p.build_stats()
    assert st["tlas_only_commits"] == 1, st
    fresh = oracle.Scene()
    fb = fresh.create_blas([(spos, sidx.reshape(-1), 1)])
    fresh.bind_tlas([fresh.create_tlas(inst1)]); assert fresh.build() == 0
    after, _ = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10, grid_width=480)
    want = fresh.trace(rays, ray_flags=0x10, n_threads=4, want_counters=False)
    _assert_parity("tlas_update_after", after, want)
    assert after.tobytes() != before.tobytes()
    got_ref, _ = _device_trace(sp.p, rays, api.TRACE_REFERENCE_ORDER, ray_flags=0x10)
    _assert_parity("tlas_update_after_reference_order", got_ref, want)
    assert sp.p.array(8).tobytes() == tri_before
    with open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "parity.jsonl"), "a") as f:
        f.write(json.dumps({"test": "tlas_only_update", "instances": int(inst1.shape[0]), "update_and_commit_ms": dt_ms}) + "\\n")
    # a different instance COUNT cannot be patched in place: still TLAS-only on the host, full upload
    sp.p.update_top_level_acceleration_structure(api.TlasHandle(t), inst1[:500])
    fresh2 = oracle.Scene(); fb2 = fresh2.create_blas([(spos, sidx.reshape(-1), 1)])
    fresh2.bind_tlas([fresh2.create_tlas(inst1[:500])]); assert fresh2.build() == 0
    got2, _ = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10, grid_width=480)
    _assert_parity("tlas_update_fewer_instances", got2, fresh2.trace(rays, ray_flags=0x10, n_threads=4, want_counters=False))
    assert sp.p.build_stats()["tlas_only_commits"] == 1
    # a new BLAS: everything is rebuilt
    b2 = sp.p.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(spos * 0.5, sidx.reshape(-1))])
    sp.p.update_top_level_acceleration_structure(api.TlasHandle(t), S.instance_grid(10, 10, b2.id, 4.0, -60.0))
    fresh3 = oracle.Scene(); fresh3.create_blas([(spos, sidx.reshape(-1), 1)]); fb3 = fresh3.create_blas([(spos * 0.5, sidx.reshape(-1), 1)])
    fresh3.bind_tlas([fresh3.create_tlas(S.instance_grid(10, 10, fb3, 4.0, -60.0))]); assert fresh3.build() == 0
    got3, _ = _device_trace(sp.p, rays, api.TRACE_AUTO, ray_flags=0x10, grid_width=480)
    _assert_parity("tlas_update_after_new_blas", got3, fresh3.trace(rays, ray_flags=0x10, n_threads=4, want_counters=False))
