"""Brute-force mesh picking on the device (rdn_pick_mesh_*) against the oracle restatement of
ray_intersect_nearest / ray_intersect_all: every topology, indexed and not, point / line tolerance, coincident primitives
(the first of equals must win), empty and one-primitive meshes, a million triangles.  Whole records, bit for bit."""
import os

import numpy as np
import pytest

import oracle
from rendiation_b200 import api, scenes as S

import helpers

pytestmark = pytest.mark.gpu
f32 = np.float32


def _rays_at(rng, targets, n, jitter):
    o = rng.uniform(-6, 6, (n, 3)).astype(f32)
    t = targets[rng.integers(0, targets.shape[0], n)] + rng.normal(0, jitter, (n, 3))
    d = (t - o).astype(np.float64)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
    return S.make_rays(o, d, 0.0, 1e30)


def _same(got, want):
    assert helpers.canonical_nan(got).tobytes() == helpers.canonical_nan(want).tobytes(), \
        (int((got["hit"] != want["hit"]).sum()), int((got["primitive_index"] != want["primitive_index"]).sum()))


@pytest.mark.parametrize("indexed", [False, True])
@pytest.mark.parametrize("topology", range(5))
def test_pick_nearest_and_all_match_the_oracle(topology, indexed):
    rng = np.random.default_rng(100 + topology * 2 + int(indexed))
    n_vert = 6000
    pos = rng.uniform(-2, 2, (n_vert, 3)).astype(f32)
    if topology in (api.TOPOLOGY_LINE_STRIP, api.TOPOLOGY_TRIANGLE_STRIP):
        pos = np.cumsum(rng.normal(0, 0.08, (n_vert, 3)), 0).astype(f32)      # a random walk: connected strips
    elif topology != api.TOPOLOGY_POINT_LIST:
        stride = 2 if topology == api.TOPOLOGY_LINE_LIST else 3                  # small primitives around random centres
        c = rng.uniform(-2, 2, (n_vert // stride, 1, 3))
        pos = (c + rng.normal(0, 0.15, (n_vert // stride, stride, 3))).reshape(-1, 3).astype(f32)
    idx = None
    if indexed:
        idx = rng.integers(0, pos.shape[0], 7001).astype(np.uint32)               # arbitrary sharing, a ragged tail (7001 % 2, % 3 != 0)
        idx[100:106] = idx[94:100]                                                # coincident primitives: ties
    else:
        pos[300:306] = pos[294:300]
    mesh = api.PickMesh(pos, idx, topology)
    count = oracle.pick_primitive_count(pos.shape[0], 0 if idx is None else idx.size, idx is not None, topology)
    assert mesh.primitive_count == count > 1000
    used = pos if idx is None else pos[idx]
    rays = _rays_at(rng, used, 700, 0.01)
    for tol, face in ((0.0, api.FACE_DOUBLE), (0.03, api.FACE_FRONT), (0.25, api.FACE_BACK)):
        want = oracle.pick_nearest(pos, idx, topology, rays, tolerance=tol, face_side=face, n_threads=os.cpu_count() or 4)
        got = mesh.ray_intersect_nearest(rays, tolerance_local=tol, triangle_face=face)
        _same(got, want)
        if tol > 0 or topology >= api.TOPOLOGY_TRIANGLE_LIST:
            assert int(want["hit"].sum()) > 20, (topology, tol)
        for k in (0, 17, 333):
            wa = oracle.pick_all(pos, idx, topology, rays[k], tolerance=tol, face_side=face)
            ga = mesh.ray_intersect_all(rays[k], tolerance_local=tol, triangle_face=face)
            _same(ga, wa)


def test_pick_ties_empty_and_tiny_meshes():
    ray = S.make_rays(np.zeros((1, 3), f32), np.array([[0, 0, 1]], f32), 0.0, 1e30)
    tri = np.array([[-1, -1, 5], [1, -1, 5], [0, 1, 5]], f32)
    pos = np.concatenate([tri + [0, 0, 2], tri, tri, tri + [0, 0, 1]]).astype(f32)
    m = api.PickMesh(pos, None, api.TOPOLOGY_TRIANGLE_LIST)
    h = m.ray_intersect_nearest(ray)[0]
    assert h["hit"] == 1 and h["primitive_index"] == 1 and h["distance"] == 5.0       # prims 1 and 2 coincide: the first stays
    assert m.ray_intersect_all(ray[0])["primitive_index"].tolist() == [0, 1, 2, 3]
    for topo, n_vert in ((api.TOPOLOGY_TRIANGLE_LIST, 2), (api.TOPOLOGY_LINE_LIST, 1), (api.TOPOLOGY_LINE_STRIP, 0), (api.TOPOLOGY_POINT_LIST, 0)):
        e = api.PickMesh(np.zeros((n_vert, 3), f32), None, topo)
        assert e.primitive_count == 0 and e.ray_intersect_nearest(ray)[0]["hit"] == 0 and e.ray_intersect_all(ray[0]).size == 0
    one = api.PickMesh(np.array([[0.05, 0, 3]], f32), None, api.TOPOLOGY_POINT_LIST)
    assert one.ray_intersect_nearest(ray, tolerance_local=0.1)[0]["hit"] == 1 and one.ray_intersect_nearest(ray, tolerance_local=0.01)[0]["hit"] == 0
    with pytest.raises(api.RdnError):
        api.PickMesh(tri, np.array([0, 1, 9], np.uint32), api.TOPOLOGY_TRIANGLE_LIST)   # vertex index out of bounds


def test_pick_a_million_triangles():
    pos, idx = S.torus_mesh(708, 708, 1.0, 0.35)
    mesh = api.PickMesh(pos, idx.reshape(-1), api.TOPOLOGY_TRIANGLE_LIST)
    assert mesh.primitive_count == 1002528
    rng = np.random.default_rng(9)
    rays = _rays_at(rng, pos, 96, 0.05)
    want = oracle.pick_nearest(pos, idx.reshape(-1), api.TOPOLOGY_TRIANGLE_LIST, rays, n_threads=os.cpu_count() or 4)
    got = mesh.ray_intersect_nearest(rays)
    _same(got, want)
    assert int(want["hit"].sum()) > 60
    wa = oracle.pick_all(pos, idx.reshape(-1), api.TOPOLOGY_TRIANGLE_LIST, rays[0])
    _same(mesh.ray_intersect_all(rays[0]), wa)
