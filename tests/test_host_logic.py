"""CPU-side checks of the product: the C-ABI library loads and exports what include/rdn_rt.h declares, the host
builder/flattener reproduce the oracle's (= the reference's) trees and arrays, and errors come back as codes."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
from rendiation_b200 import api, scenes as S

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rdn_rt.h")).read()
    declared = set(re.findall(r"\b(rdn_(?:rt|bvh|pick|sbt)_[a-z0-9_]+)\s*\(", header))
    assert declared == set(api.EXPORTED_SYMBOLS), declared ^ set(api.EXPORTED_SYMBOLS)
    L = ctypes.CDLL(api.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert b"sm_100a" in api.lib().rdn_rt_version()


def test_record_sizes_match_header():
    assert api.HIT_DTYPE.itemsize == 32 and S.RAY_DTYPE.itemsize == 32 and api.MESH_HIT_DTYPE.itemsize == 32
    assert S.INSTANCE_DTYPE.itemsize == 84 and api.FLAT_BVH_NODE_DTYPE.itemsize == 64
    assert api.DEV_NODE_DTYPE.itemsize == 48 and api.TRI_RECORD_DTYPE.itemsize == 64 and api.WIDE_NODE_DTYPE.itemsize == 64
    assert oracle.HIT_DTYPE == api.HIT_DTYPE


def _eq(a, b):
    """numeric equality field by field (+0 == -0: f32::min/max leave the sign of zero unspecified)"""
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.dtype.names:
        return all(_eq(a[f], b[f]) for f in a.dtype.names)
    return bool(np.array_equal(a, b))


@pytest.mark.parametrize("opt", [(50, 2), (10, 50), (15, 10), (3, 1)])
@pytest.mark.parametrize("strategy", ["sah", "balance"])
def test_flatten_bvh_builder_matches_oracle(strategy, opt):
    pos, idx = S.torus_mesh(40, 24)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([pos[tri].min(1), pos[tri].max(1)], 1)
    ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH if strategy == "sah" else oracle.STRATEGY_BALANCE, 4, *opt)
    pb = api.FlattenBVH(boxes, api.SAH(4) if strategy == "sah" else api.BalanceTree(), api.TreeBuildOption(*opt))
    on, pn = ob.nodes, pb.nodes
    assert _eq(on, pn)
    assert np.array_equal(ob.sorted_primitive_index, pb.sorted_primitive_index)


@pytest.mark.parametrize("strategy", ["sah", "balance"])
def test_parallel_builder_is_the_sequential_tree(strategy, monkeypatch):
    """above 32,768 primitives the product builds subtrees on worker threads and splices them in pre-order: node for node the
    oracle's (= the reference's single-threaded) tree, for any thread count"""
    pos, idx = S.torus_mesh(170, 130)
    tri = idx.reshape(-1, 3)
    assert tri.shape[0] > (1 << 15)
    boxes = np.concatenate([pos[tri].min(1), pos[tri].max(1)], 1)
    ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH if strategy == "sah" else oracle.STRATEGY_BALANCE, 4, 50, 2)
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("RDN_BUILD_THREADS", threads)
        pb = api.FlattenBVH(boxes, api.SAH(4) if strategy == "sah" else api.BalanceTree(), api.TreeBuildOption(50, 2))
        assert _eq(ob.nodes, pb.nodes), threads
        assert np.array_equal(ob.sorted_primitive_index, pb.sorted_primitive_index), threads


def test_flattened_scene_does_not_depend_on_the_thread_count(monkeypatch):
    """the multi-threaded build / flatten is deterministic: every array of the blob is byte-identical for 1, 3 and all threads"""
    pos, idx = S.torus_mesh(200, 130)   # 52,000 triangles: above the parallel thresholds
    blobs = []
    for threads in ("1", "3", None):
        if threads is None:
            monkeypatch.delenv("RDN_BUILD_THREADS", raising=False)
        else:
            monkeypatch.setenv("RDN_BUILD_THREADS", threads)
        sp = helpers.single_mesh_scene(pos, idx, S.mat4_translate(0, 0, -10), devices=(), product=True)
        blobs.append({k: v.tobytes() for k, v in sp.p.arrays().items()})
    for other in blobs[1:]:
        assert other.keys() == blobs[0].keys()
        for k in other:
            assert other[k] == blobs[0][k], k


def test_builder_edge_cases():
    for n in (0, 1, 2, 3):
        boxes = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (n, 1))
        ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, 50, 2)
        pb = api.FlattenBVH(boxes, api.SAH(4), api.TreeBuildOption(50, 2))
        assert _eq(ob.nodes, pb.nodes) and np.array_equal(ob.sorted_primitive_index, pb.sorted_primitive_index)
    # identical centres -> SAH degenerate -> BalanceTree fallback
    boxes = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (9, 1))
    ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, 50, 2)
    assert ob.balance_fallbacks > 0
    pb = api.FlattenBVH(boxes, api.SAH(4), api.TreeBuildOption(50, 2))
    assert _eq(ob.nodes, pb.nodes)


def test_build_bvh_for_abstract_mesh_matches_box_build():
    pos, idx = S.uv_sphere_mesh(12, 10)
    tri = idx.reshape(-1, 3)
    boxes = np.concatenate([pos[tri].min(1), pos[tri].max(1)], 1)
    a = api.build_bvh_for_abstract_mesh(pos, idx, api.SAH(4), api.TreeBuildOption(50, 2))
    b = api.FlattenBVH(boxes, api.SAH(4), api.TreeBuildOption(50, 2))
    assert _eq(a.nodes, b.nodes)


def _flat_pair(builder):
    sp = builder(devices=(), product=True)
    sp = sp[0] if isinstance(sp, tuple) else sp
    return sp.o.view(), sp.p.arrays(), sp


def test_flattened_scene_matches_oracle_reference_fixture():
    ov, pa, sp = _flat_pair(helpers.reference_fixture)
    assert np.array_equal(pa["tlas_binding"], ov["tlas_binding"])
    assert np.array_equal(pa["tlas_root"][:, 0], ov["tlas_bvh_root"])
    assert _eq(pa["tlas_bvh_forest"], ov["tlas_bvh_forest"])
    assert _eq(pa["tlas_bounding"], ov["tlas_bounding"])
    assert _eq(pa["tri_bvh_forest"], ov["tri_bvh_forest"])
    assert np.array_equal(pa["blas_meta"][:, :2], ov["blas_meta_info"])
    for f in ("bvh_root_idx", "geometry_idx", "primitive_start", "geometry_flags"):
        assert np.array_equal(pa["geometry_meta"][f], ov["tri_bvh_root"][f])
    for f, g in (("transform_inv", "transform_inv"), ("instance_custom_index", "instance_custom_index"), ("sbt_offset", "sbt_offset"),
                 ("flags", "flags"), ("blas", "blas")):
        assert np.array_equal(pa["instances"][f], ov["tlas_data"][g])
    # slot_info = indices_redirect - primitive_start of the owning geometry
    starts = ov["tri_bvh_root"]["primitive_start"]
    owner = np.searchsorted(starts, ov["indices_redirect"], side="right") - 1
    assert np.array_equal(pa["slot_info"][:, 0], ov["indices_redirect"] - starts[owner])
    # pre-gathered triangle records carry the reference's vertices in slot order
    tri = ov["indices"].reshape(-1, 3)[ov["indices_redirect"]]
    v0, v1, v2 = ov["vertices"][tri[:, 0]], ov["vertices"][tri[:, 1]], ov["vertices"][tri[:, 2]]
    assert np.array_equal(pa["triangles"]["v0"], v0)
    assert np.array_equal(pa["triangles"]["e1"], v1 - v0) and np.array_equal(pa["triangles"]["e2"], v2 - v0)


def _exponential_scene(devices=(), product=True):
    """200 triangles whose size and position halve from one to the next: the SAH split peels a few off per level, depth 50 is
    reached with a hundred left, and that leaf (more than 16 slots) becomes a chain of wide nodes"""
    x = 2.0 ** (60 - np.arange(200)).astype(np.float64)
    tri = np.zeros((200, 3, 3))
    tri[:, 0] = np.stack([x, 0 * x, 0 * x - 5], 1)
    tri[:, 1] = np.stack([x * 1.05, 0 * x, 0 * x - 5], 1)
    tri[:, 2] = np.stack([x, x * 0.05, 0 * x - 5], 1)
    sp = helpers.ScenePair(devices, product)
    b = sp.blas([(tri.reshape(-1, 3).astype(np.float32), None, 1)])
    sp.bind([sp.tlas(S.make_instance(S.mat4_identity(), b))])
    return sp.build()


@pytest.mark.parametrize("builder", [helpers.sphere_c1, _exponential_scene])
def test_wide_nodes_cover_the_reference_tree(builder):
    """every inner reference node appears once with its children's exact boxes; leaves decode to the same slot ranges"""
    ov, pa, sp = _flat_pair(builder)
    wide, forest, gm = pa["wide_nodes"], pa["tri_bvh_forest"], pa["geometry_meta"][0]
    LEAF = 0x80000000
    seen_slots = []

    def chain_slots(ref, rn):
        """slots behind a reference that stands for reference leaf rn: one leaf reference, or (leaves of more than 16 slots) a
        chain of nodes that repeat the leaf's box"""
        if ref & LEAF:
            start, cnt = ref & ((1 << 27) - 1), ((ref >> 27) & 15) + 1
            return list(range(start, start + cnt))
        w = wide[ref]
        out = []
        for cmin, cmax, r in ((w["c0_min"], w["c0_max"], w["ref0"]), (w["c1_min"], w["c1_max"], w["ref1"])):
            assert np.array_equal(cmin, rn["aabb_min"]) and np.array_equal(cmax, rn["aabb_max"])
            out += chain_slots(int(r), rn)
        return out

    def walk(ref, ref_node):
        rn = forest[ref_node]
        if rn["hit_next"] == rn["miss_next"]:
            slots = chain_slots(ref, rn)
            assert slots == list(range(rn["range"][0], rn["range"][1]))
            seen_slots.extend(slots)
            return
        w = wide[ref]
        left, right = ref_node + 1, None
        assert rn["hit_next"] == left
        # right child = the miss link of the left child
        right = int(forest[left]["miss_next"])
        for cmin, cmax, child in ((w["c0_min"], w["c0_max"], left), (w["c1_min"], w["c1_max"], right)):
            assert np.array_equal(cmin, forest[child]["aabb_min"]) and np.array_equal(cmax, forest[child]["aabb_max"])
        walk(int(w["ref0"]), left)
        walk(int(w["ref1"]), right)

    import sys
    sys.setrecursionlimit(10000)
    root = wide[gm["wide_root"]]
    assert np.array_equal(root["c0_min"], forest[0]["aabb_min"]) and np.isnan(root["c1_min"]).all() and root["ref1"] == 0x7FFFFFFE
    walk(int(root["ref0"]), 0)
    assert sorted(seen_slots) == list(range(len(pa["triangles"])))


@pytest.mark.parametrize("builder", [helpers.sphere_c1, helpers.reference_fixture, _exponential_scene])
def test_wide4_nodes_cover_the_reference_tree(builder):
    """the 4-wide view: every node holds the exact boxes of the reference nodes it stands for (grandchildren of an inner node, a
    leaf child kept as it is), unused slots are NaN / REF_EMPTY, and the leaves decode to every slot exactly once"""
    os.environ["RDN_ORDERED_VARIANT"] = "60"  # the view is only emitted for the experiment that walks it
    try:
        ov, pa, sp = _flat_pair(builder)
    finally:
        del os.environ["RDN_ORDERED_VARIANT"]
    w4, LEAF, EMPTY = pa["wide4_nodes"], 0x80000000, 0x7FFFFFFE

    def check_tree(forest, root_node, w4_root, slot_lo, slot_hi):
        is_leaf = lambda k: forest[k]["hit_next"] == forest[k]["miss_next"]
        kids_of = lambda k: (k + 1, int(forest[k + 1]["miss_next"]))
        seen = []

        def chain_slots(ref, k):
            if ref & LEAF:
                start, cnt = ref & ((1 << 27) - 1), ((ref >> 27) & 15) + 1
                return list(range(start, start + cnt))
            out = []
            for c in w4[ref]["child"]:
                if c["ref"] == EMPTY:
                    continue
                assert np.array_equal(c["bmin"], forest[k]["aabb_min"]) and np.array_equal(c["bmax"], forest[k]["aabb_max"])
                out += chain_slots(int(c["ref"]), k)
            return out

        def leaf(ref, k):
            slots = chain_slots(ref, k)
            assert slots == list(range(forest[k]["range"][0], forest[k]["range"][1]))
            seen.extend(slots)

        def visit(ref, k):  # ref stands for reference node k
            if is_leaf(k):
                return leaf(ref, k)
            node = w4[ref]["child"]
            want = []
            for c in kids_of(k):
                want += [c] if is_leaf(c) else list(kids_of(c))
            for j, g in enumerate(want):
                assert np.array_equal(node[j]["bmin"], forest[g]["aabb_min"]) and np.array_equal(node[j]["bmax"], forest[g]["aabb_max"])
                visit(int(node[j]["ref"]), g)
            for j in range(len(want), 4):
                assert node[j]["ref"] == EMPTY and np.isnan(node[j]["bmin"]).all()

        root = w4[w4_root]["child"]
        assert np.array_equal(root[0]["bmin"], forest[root_node]["aabb_min"]) and all(root[j]["ref"] == EMPTY for j in (1, 2, 3))
        visit(int(root[0]["ref"]), root_node)
        assert sorted(seen) == list(range(slot_lo, slot_hi))

    import sys
    sys.setrecursionlimit(20000)
    tri_forest, slot = pa["tri_bvh_forest"], 0
    for g in pa["geometry_meta"]:
        n_slots = int(tri_forest[g["bvh_root_idx"]]["range"][1] - tri_forest[g["bvh_root_idx"]]["range"][0])
        check_tree(tri_forest, int(g["bvh_root_idx"]), int(g["wide4_root"]), slot, slot + n_slots)
        slot += n_slots
    tlas_forest = pa["tlas_bvh_forest"]
    for t in pa["tlas_root"]:
        r = int(t[0])
        check_tree(tlas_forest, r, int(t[7]), int(tlas_forest[r]["range"][0]), int(tlas_forest[r]["range"][1]))


def test_errors_are_status_codes_not_aborts():
    s = api.NaiveSahBVHSystem(devices=())
    b = s.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(S.CUBE_POSITION, S.CUBE_INDEX)])
    t = s.create_top_level_acceleration_structure(S.make_instance(S.mat4_identity(), b.id))
    s.bind_tlas([t])
    s.commit()
    assert s.bind_tlas_max_len() == 0xFFFFFFFF
    s.delete_bottom_level_acceleration_structure(b)
    with pytest.raises(api.RdnError) as e:  # the reference panics on unwrap() (naive/mod.rs:273-275)
        s.commit()
    assert e.value.code == -3
    with pytest.raises(api.RdnError):
        s.delete_top_level_acceleration_structure(api.TlasHandle(99))
    bad = api.BottomLevelAccelerationStructureBuildSource(S.CUBE_POSITION, np.array([0, 1, 999], np.uint32))
    s2 = api.NaiveSahBVHSystem(devices=())
    s2.create_bottom_level_acceleration_structure([bad])
    with pytest.raises(api.RdnError) as e:
        s2.commit()
    assert e.value.code == -4


def test_no_cpu_fallback():
    """a host-only scene can build and flatten but must refuse to trace: the product has no CPU path"""
    s = api.NaiveSahBVHSystem(devices=())
    b = s.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(S.CUBE_POSITION, S.CUBE_INDEX)])
    t = s.create_top_level_acceleration_structure(S.make_instance(S.mat4_translate(0, 0, -3), b.id))
    s.bind_tlas([t])
    with pytest.raises(api.RdnError):
        s.trace_closest_batch(S.pinhole_rays(4, 4))
    with pytest.raises(api.RdnError):
        s.compact_u32(np.arange(4, dtype=np.uint32), np.ones(4, np.uint8))
    # and the package never imports the oracle
    import rendiation_b200
    src_dir = os.path.dirname(rendiation_b200.__file__)
    for dirpath, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "oracle/" not in txt and "liboracle" not in txt, f


def test_tlas_wide_view_reaches_every_instance_through_single_slot_leaves():
    """TLAS leaves of several instances are emitted as subtrees over contiguous halves of their slots: every instance slot of
    every TLAS is reachable exactly once, as a one-slot leaf, and every box on the way down contains the exact instance boxes
    below it (nested boxes keep the slab test monotone)"""
    sp, handles = helpers.reference_fixture(devices=(), product=True)
    pa = sp.p.arrays()
    wide, roots, tb = pa["wide_nodes"], pa["tlas_root"], pa["tlas_bounding"]
    LEAF, SPECIAL = 0x80000000, 0x7F000000
    total = 0
    for r in roots:
        wide_root = int(r[1])
        if wide_root == 0x7FFFFFFE:
            continue
        seen = []

        def walk(ref, bmin, bmax):
            if ref & LEAF:
                start, cnt = ref & ((1 << 27) - 1), ((ref >> 27) & 15) + 1
                assert cnt == 1, "a multi-slot instance leaf survived"
                seen.append(start)
                assert (tb[start]["world_min"] >= bmin).all() and (tb[start]["world_max"] <= bmax).all()
                return
            assert ref < SPECIAL
            w = wide[ref]
            for cmin, cmax, child in ((w["c0_min"], w["c0_max"], int(w["ref0"])), (w["c1_min"], w["c1_max"], int(w["ref1"]))):
                if child == 0x7FFFFFFE:
                    continue
                assert (cmin >= bmin).all() and (cmax <= bmax).all()
                walk(child, cmin, cmax)

        root = wide[wide_root]
        walk(int(root["ref0"]), root["c0_min"], root["c0_max"])
        assert len(seen) == len(set(seen)) and seen, seen
        total += len(seen)
    assert total == len(tb)


# ---- worker pool of the host builder / flattener (bvh_builder.cpp) -----------------------------------------------------------------
def _instanced_host_scene(n_side=70, spacing=3.5):
    pos, idx = S.uv_sphere_mesh(24, 24)
    s = api.NaiveSahBVHSystem(devices=())
    b = s.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
    t = s.create_top_level_acceleration_structure(S.instance_grid(n_side, n_side, b.id, spacing, -200.0))
    s.bind_tlas([t])
    s.commit()
    return s, b, t


def _arrays(s):
    return [s.array(aid).tobytes() for aid in range(len(api.ARRAYS))]


def test_pool_sections_give_the_sequential_arrays():
    """4,900 instances: the TLAS build takes the parallel path (top splits by all threads, subtrees on the pool); the flattened arrays
    must be those of a one-thread build, and a TLAS-only update must give what a fresh scene with the moved instances gives"""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import hashlib, test_host_logic as T\n"
            "s, b, t = T._instanced_host_scene()\n"
            "print(hashlib.sha256(b''.join(T._arrays(s))).hexdigest())\n") % (ROOT, os.path.join(ROOT, "tests"))
    digests = set()
    for env in ({"RDN_BUILD_THREADS": "1"}, {"RDN_BUILD_POOL": "0"}, {}, {"RDN_BUILD_THREADS": "3"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        digests.add(r.stdout.strip().splitlines()[-1])
    assert len(digests) == 1, digests
    s, b, t = _instanced_host_scene()
    moved = S.instance_grid(70, 70, b.id, 3.6, -190.0)
    s.update_top_level_acceleration_structure(t, moved)
    s.commit()
    fresh = api.NaiveSahBVHSystem(devices=())
    pos, idx = S.uv_sphere_mesh(24, 24)
    fb = fresh.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
    fresh.bind_tlas([fresh.create_top_level_acceleration_structure(S.instance_grid(70, 70, fb.id, 3.6, -190.0))])
    fresh.commit()
    assert _arrays(s) == _arrays(fresh)


def test_concurrent_commits_share_the_pool():
    """commits of different scenes from different host threads: their parallel sections queue on the one pool"""
    import threading
    want = _arrays(_instanced_host_scene()[0])
    got, errors = [None] * 4, []

    def work(k):
        try:
            for _ in range(3):
                got[k] = _arrays(_instanced_host_scene()[0])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [th.start() for th in threads]
    [th.join(300) for th in threads]
    assert not errors and all(g == want for g in got)


def test_a_forked_child_gets_its_own_pool():
    """the pool's threads do not exist in a forked child: the first section there must start a fresh pool, not wait for ghosts"""
    want = _arrays(_instanced_host_scene()[0])  # (the parent's pool is running now)
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        code = 1
        try:
            os.close(r)
            ok = _arrays(_instanced_host_scene()[0]) == want
            os.write(w, b"ok" if ok else b"differs")
            code = 0
        finally:
            os._exit(code)
    os.close(w)
    import select
    ready, _, _ = select.select([r], [], [], 120)
    msg = os.read(r, 16) if ready else b"timeout"
    os.close(r)
    if not ready:
        os.kill(pid, 9)
    os.waitpid(pid, 0)
    assert msg == b"ok", msg
