"""Known answers for the oracle's picking restatement (oracle_pick.c): Ray3 x Point / LineSegment with tolerance
(math/geometry/src/dimension3/intersection.rs:79-121, ray3.rs:48-145), primitive counting per topology
(content/mesh/core/src/container/attributes/access.rs:142-150) and the nearest / all queries
(content/mesh/core/src/feature/intersection.rs:3-37).  CPU only."""
import numpy as np

import oracle
from rendiation_b200 import scenes as S

f32 = np.float32
PT, LL, LS, TL, TS = range(5)


def _ray(o, d):
    d = np.asarray(d, np.float64)
    return S.make_rays(np.asarray([o], f32), (d / np.linalg.norm(d)).astype(f32)[None, :], 0.0, 1e30)


def test_primitive_count_per_topology():
    for count, want in ((7, (7, 3, 6, 2, 5)), (0, (0, 0, 0, 0, 0)), (1, (1, 0, 0, 0, 0)), (2, (2, 1, 1, 0, 0)), (3, (3, 1, 2, 1, 1)), (6, (6, 3, 5, 2, 4))):
        for topo in range(5):
            assert oracle.pick_primitive_count(count, 0, False, topo) == want[topo], (count, topo)
            assert oracle.pick_primitive_count(99, count, True, topo) == want[topo], (count, topo)


def test_point_tolerance_and_backward_rejection():
    ray = _ray((0, 0, 0), (0, 0, 1))
    pts = np.array([[0.1, 0.0, 5.0]], f32)
    h = oracle.pick_nearest(pts, None, PT, ray, tolerance=0.2)[0]
    assert h["hit"] == 1 and h["primitive_index"] == 0 and (h["px"], h["py"], h["pz"]) == (f32(0.1), 0.0, 5.0)
    assert h["distance"] == np.sqrt(f32(f32(0.1) * f32(0.1)) + f32(25.0), dtype=f32)
    assert oracle.pick_nearest(pts, None, PT, ray, tolerance=0.05)[0]["hit"] == 0
    exact = np.array([[0.5, 0.0, 4.0]], f32)                                                # dist_sq = 16.25 - 16 = 0.25 exactly
    assert oracle.pick_nearest(exact, None, PT, ray, tolerance=0.5)[0]["hit"] == 1         # dist_sq == t*t is a hit (`>` rejects)
    assert oracle.pick_nearest(exact, None, PT, ray, tolerance=0.4999)[0]["hit"] == 0
    behind = np.array([[0.0, 0.0, -1.0]], f32)
    assert oracle.pick_nearest(behind, None, PT, ray, tolerance=10.0)[0]["hit"] == 0      # tca < 0: never, whatever the tolerance


def test_segment_regions():
    ray = _ray((0, 0, 0), (0, 0, 1))
    seg = lambda a, b: np.array([a, b], f32)
    # region 0: closest points interior to both; closest point on the RAY is reported, distance along the ray
    h = oracle.pick_nearest(seg((-1, 0.3, 4), (1, 0.3, 4)), None, LL, ray, tolerance=0.31)[0]
    assert h["hit"] == 1 and abs(h["distance"] - 4.0) < 1e-6 and abs(h["px"]) < 1e-6 and abs(h["py"]) < 1e-6 and abs(h["pz"] - 4.0) < 1e-6
    assert oracle.pick_nearest(seg((-1, 0.3, 4), (1, 0.3, 4)), None, LL, ray, tolerance=0.29)[0]["hit"] == 0
    # regions 1 / 5: the segment ends before reaching the ray; the nearer end point decides
    for a, b in (((1, 0, 4), (2, 0, 4)), ((2, 0, 4), (1, 0, 4))):
        assert oracle.pick_nearest(seg(a, b), None, LL, ray, tolerance=1.001)[0]["hit"] == 1
        assert oracle.pick_nearest(seg(a, b), None, LL, ray, tolerance=0.999)[0]["hit"] == 0
    # regions 2 / 3 / 4: the closest point on the ray would be behind the origin -> clamped to the origin
    h = oracle.pick_nearest(seg((-1, 0.5, -2), (1, 0.5, -2)), None, LL, ray, tolerance=2.1)[0]
    assert h["hit"] == 1 and h["distance"] == 0.0 and (h["px"], h["py"], h["pz"]) == (0.0, 0.0, 0.0)
    assert oracle.pick_nearest(seg((-1, 0.5, -2), (1, 0.5, -2)), None, LL, ray, tolerance=2.0)[0]["hit"] == 0   # sqrt(4.25) > 2
    # parallel
    h = oracle.pick_nearest(seg((0.5, 0, 1), (0.5, 0, 3)), None, LL, ray, tolerance=0.5)[0]
    assert h["hit"] == 1 and abs(h["distance"] - 3.0) < 1e-6 or abs(h["distance"] - 1.0) < 1e-6
    assert oracle.pick_nearest(seg((0.5, 0, 1), (0.5, 0, 3)), None, LL, ray, tolerance=0.49)[0]["hit"] == 0


def test_segment_distance_against_float64_clamping():
    """sq distance decision vs an independent float64 closest-point computation, away from the threshold"""
    rng = np.random.default_rng(3)
    n = 4000
    a = rng.uniform(-2, 2, (n, 3)); b = a + rng.normal(0, 0.7, (n, 3))
    o = rng.uniform(-3, 3, 3); d = rng.normal(size=3); d /= np.linalg.norm(d)
    ray = _ray(o, d)
    o, d = np.array([ray["ox"][0], ray["oy"][0], ray["oz"][0]], np.float64), np.array([ray["dx"][0], ray["dy"][0], ray["dz"][0]], np.float64)
    a32, b32 = a.astype(f32), b.astype(f32)
    a, b = a32.astype(np.float64), b32.astype(np.float64)
    # brute: minimise |o + s d - (a + u (b - a))| over s >= 0, u in [0, 1] by alternating projection (convex, converges)
    u = np.full(n, 0.5); s = np.zeros(n)
    for _ in range(200):
        p = a + u[:, None] * (b - a)
        s = np.maximum(0.0, ((p - o) * d).sum(1))
        q = o + s[:, None] * d
        u = np.clip(((q - a) * (b - a)).sum(1) / np.maximum(((b - a) ** 2).sum(1), 1e-300), 0.0, 1.0)
    dist = np.linalg.norm(o + s[:, None] * d - (a + u[:, None] * (b - a)), axis=1)
    tol = 0.8
    pos = np.stack([a32, b32], 1).reshape(-1, 3)
    for k in np.nonzero(np.abs(dist - tol) > 1e-3)[0][:1500]:
        h = oracle.pick_nearest(pos[2 * k:2 * k + 2], None, LL, ray, tolerance=tol)[0]
        assert bool(h["hit"]) == bool(dist[k] < tol), (k, dist[k])
        cos = abs(((b[k] - a[k]) * d).sum()) / np.linalg.norm(b[k] - a[k])
        if h["hit"] and cos < 0.9:  # (alternating projection converges slowly for nearly parallel pairs)
            assert abs(h["distance"] - s[k]) < 2e-3 * max(1.0, s[k])


def test_nearest_is_the_first_of_equals_and_all_lists_in_primitive_order():
    tri = np.array([[-1, -1, 5], [1, -1, 5], [0, 1, 5]], f32)
    pos = np.concatenate([tri + [0, 0, 2], tri, tri, tri + [0, 0, 1]]).astype(f32)   # prims 1 and 2 coincide, 0 and 3 are behind them
    ray = _ray((0, 0, 0), (0, 0, 1))
    h = oracle.pick_nearest(pos, None, TL, ray)[0]
    assert h["hit"] == 1 and h["primitive_index"] == 1 and h["distance"] == 5.0
    allh = oracle.pick_all(pos, None, TL, ray[0])
    assert allh["primitive_index"].tolist() == [0, 1, 2, 3] and allh["distance"].tolist() == [7.0, 5.0, 5.0, 6.0]
    # strips share vertices: 4 points -> 3 segments / 2 triangles
    strip = np.array([[-1, 0.1, 3], [1, 0.1, 3], [1, 0.1, 6], [-1, 0.1, 6]], f32)
    assert oracle.pick_all(strip, None, LS, ray[0], tolerance=0.2)["primitive_index"].tolist() == [0, 2]
    quad = np.array([[-1, -1, 5], [1, -1, 5], [-1, 1, 5], [1, 1, 5]], f32)   # strip triangles (0,1,2) and (1,2,3)
    idx = np.array([0, 1, 2, 3], np.uint32)
    assert oracle.pick_all(quad, idx, TS, _ray((-0.5, -0.5, 0), (0, 0, 1))[0])["primitive_index"].tolist() == [0]
    assert oracle.pick_all(quad, idx, TS, _ray((0.5, 0.5, 0), (0, 0, 1))[0])["primitive_index"].tolist() == [1]
    assert oracle.pick_all(quad, None, TL, _ray((0.5, 0.5, 0), (0, 0, 1))[0]).size == 0                  # list: only (0,1,2)
