"""The device SAH build (rdn_bvh_build_device, csrc/build_device.cu) gives the host builder's — i.e. the
reference's — tree node for node: boxes, ranges, pre-order numbering, split axes, and the sorted primitive index."""
import numpy as np
import pytest

import oracle
from rendiation_b200 import api, scenes as S

pytestmark = pytest.mark.gpu
f32 = np.float32


def _boxes(pos, idx):
    tri = idx.reshape(-1, 3)
    return np.concatenate([pos[tri].min(1), pos[tri].max(1)], 1).astype(f32)


def _same_tree(dev, ob):
    dn, on = dev.nodes, ob.nodes
    assert dn.shape == on.shape
    assert np.array_equal(dev.sorted_primitive_index, ob.sorted_primitive_index)
    assert np.array_equal(dn["bmin"], on["bmin"]) and np.array_equal(dn["bmax"], on["bmax"])
    for f in ("start", "end", "self_index", "has_child"):
        assert np.array_equal(dn[f], on[f]), f
    inner = on["has_child"] != 0
    assert np.array_equal(dn["left_count"][inner], on["left_count"][inner]) and np.array_equal(dn["split_axis"][inner], on["split_axis"][inner])


@pytest.mark.parametrize("opt", [(50, 2), (10, 50), (3, 1)])
def test_device_build_is_the_reference_tree(opt):
    rng = np.random.default_rng(4)
    c = rng.uniform(-1, 1, (5000, 3)).astype(f32)
    r = rng.uniform(0.01, 0.3, (5000, 3)).astype(f32)
    for boxes in (_boxes(*S.torus_mesh(170, 130)), _boxes(*S.uv_sphere_mesh(40, 24)), np.concatenate([c - r, c + r], 1),
                  np.tile(np.array([[0, 0, 0, 1, 1, 1]], f32), (9, 1)), np.zeros((1, 6), f32)):
        dev = api.FlattenBVH(boxes, api.SAH(4), api.TreeBuildOption(*opt), device=0)
        _same_tree(dev, oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, *opt))


def test_device_build_of_a_million_triangles_and_its_fallbacks():
    boxes = _boxes(*S.torus_mesh(708, 708))
    dev = api.FlattenBVH(boxes, api.SAH(4), api.TreeBuildOption(50, 2), device=0)
    assert dev.built_on_device
    host = api.FlattenBVH(boxes, api.SAH(4), api.TreeBuildOption(50, 2))
    assert dev.nodes.tobytes() == host.nodes.tobytes() and np.array_equal(dev.sorted_primitive_index, host.sorted_primitive_index)
    # more buckets than the device build covers, or a long run of identical centres: silently the host builder, same tree
    eight = api.FlattenBVH(boxes[:20000], api.SAH(8), api.TreeBuildOption(50, 2), device=0)
    assert not eight.built_on_device
    _same_tree(eight, oracle.FlattenBVH(boxes[:20000], oracle.STRATEGY_SAH, 8, 50, 2))
    many = np.tile(np.array([[0, 0, 0, 1, 1, 1]], f32), (300, 1))
    same_centres = api.FlattenBVH(many, api.SAH(4), api.TreeBuildOption(50, 2), device=0)
    assert not same_centres.built_on_device
    _same_tree(same_centres, oracle.FlattenBVH(many, oracle.STRATEGY_SAH, 4, 50, 2))


def test_commit_with_the_device_builder_flattens_the_same_scene(monkeypatch):
    """RDN_COMMIT_DEVICE_BUILD=1: geometry trees above the size threshold come from the device SAH builder; the flattened scene
    (every array of the blob) and the traced records are those of the host-built scene"""
    from rendiation_b200 import scenes as S
    import helpers

    def build():
        sp = helpers.ScenePair(product=True)
        tpos, tidx = S.torus_mesh(96, 64, 1.0, 0.35)           # 12,288 triangles: device
        spos, sidx = S.uv_sphere_mesh(12, 8)                   # small: stays on the host
        a = sp.blas([(tpos, tidx.reshape(-1), 1)])
        b = sp.blas([(spos, sidx.reshape(-1), 1)])
        T = S.mat4_translate
        sp.bind([sp.tlas(np.concatenate([S.make_instance(S.mat4_mul(T(0, 0, -8), S.mat4_rotate_x(-0.6)), a),
                                         S.make_instance(T(2.5, 1.5, -7), b, custom_index=3)]))])
        return sp.build()

    host = build()
    assert host.p.build_stats()["device_built_trees"] == 0
    monkeypatch.setenv("RDN_COMMIT_DEVICE_BUILD", "1")
    monkeypatch.setenv("RDN_COMMIT_DEVICE_BUILD_MIN", "4096")
    dev = build()
    assert dev.p.build_stats()["device_built_trees"] == 1
    ha, da = host.p.arrays(), dev.p.arrays()
    assert ha.keys() == da.keys()
    for name in ha:
        assert ha[name].tobytes() == da[name].tobytes(), name
    rays = S.pinhole_rays(96, 64, 0.01, 100.0)
    got = dev.p.trace_closest_batch(rays, ray_flags=0x10, grid_width=96)
    want, _ = dev.o.trace(rays, ray_flags=0x10, n_threads=4)
    assert got.tobytes() == want.tobytes() and int((got["instance_id"] != 0xFFFFFFFF).sum()) > 100
