"""tools/level_sync_build.py — the level-synchronous, data-parallel formulation of the reference's SAH builder that a CUDA build
 will follow pass by pass — must give the sequential builder's tree node for node."""
import os
import sys

import numpy as np
import pytest

import oracle
from rendiation_b200 import scenes as S

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import level_sync_build  # noqa: E402


def _check(boxes, opt):
    ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, *opt)
    nodes, order = level_sync_build.build(boxes, 4, *opt)
    on = ob.nodes
    assert on.shape[0] == nodes["start"].shape[0]
    assert np.array_equal(ob.sorted_primitive_index, order)
    assert np.array_equal(on["bmin"], nodes["bmin"]) and np.array_equal(on["bmax"], nodes["bmax"])
    for f in ("start", "end", "self_index", "has_child"):
        assert np.array_equal(on[f], nodes[f]), f
    inner = on["has_child"] != 0
    assert np.array_equal(on["left_count"][inner], nodes["left_count"][inner]) and np.array_equal(on["split_axis"][inner], nodes["split_axis"][inner])


@pytest.mark.parametrize("opt", [(50, 2), (10, 50), (3, 1)])
def test_level_synchronous_build_equals_the_recursive_one(opt):
    for pos, idx in (S.torus_mesh(48, 30), S.uv_sphere_mesh(40, 24)):   # the sphere's pole fans exercise the BalanceTree fallback
        tri = idx.reshape(-1, 3)
        _check(np.concatenate([pos[tri].min(1), pos[tri].max(1)], 1), opt)


def test_level_synchronous_build_edge_cases():
    rng = np.random.default_rng(2)
    c = rng.uniform(-1, 1, (700, 3)).astype(np.float32)
    r = rng.uniform(0.01, 0.4, (700, 3)).astype(np.float32)
    _check(np.concatenate([c - r, c + r], 1), (50, 2))                                        # overlapping random boxes
    for n in (1, 2, 3, 9):
        _check(np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (n, 1)), (50, 2))          # identical boxes: every split degenerate
    flat = np.concatenate([c * [1, 0, 1], c * [1, 0, 1]], 1).astype(np.float32)
    _check(flat, (50, 2))                                                                     # zero-extent boxes in a plane
