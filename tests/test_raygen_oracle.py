"""Known answers and invariants of the ray-generation restatement (oracle/raygen.py), CPU only."""
import numpy as np

from oracle import raygen as R
from rendiation_b200 import scenes as S

f32 = np.float32


def test_radical_inverse_and_hammersley_known_answers():
    # radical_inverse_vdc mirrors the bits of the index around the binary point (sampling.rs:43-52)
    got = R.radical_inverse_vdc(np.array([0, 1, 2, 3, 4, 5, 255], np.uint32))
    want = np.array([0.0, 0.5, 0.25, 0.75, 0.125, 0.625, 255 / 256], f32)
    assert np.allclose(got, want, atol=1e-7)
    h = R.hammersley_2d(np.array([0, 64, 128], np.uint32), 256)
    assert np.array_equal(h[:, 0], np.array([0, 0.25, 0.5], f32)) and h[0, 1] == 0 and abs(h[2, 1] - 2.0 ** -8) < 1e-9


def test_pcg_and_xxhash_are_deterministic_and_in_range():
    px = np.arange(64, dtype=np.uint32)
    s = R.xxhash32(px, px[::-1].copy(), np.full(64, 7, np.uint32))
    assert s.dtype == np.uint32 and len(set(s.tolist())) == 64
    # scalar re-statement of the hash for one value (u32 wrap-around arithmetic)
    M = 0xFFFFFFFF
    p0, p1, p2, p3 = 2246822519, 3266489917, 668265263, 374761393
    x, y, z = 5, 58, 7
    h = (z + p3 + x * p1) & M
    h = (p2 * (((h << 17) | (h >> 15)) & M)) & M
    h = (h + y * p1) & M
    h = (p2 * (((h << 17) | (h >> 15)) & M)) & M
    h = (p0 * (h ^ (h >> 15))) & M
    h = (p1 * (h ^ (h >> 13))) & M
    assert int(s[5]) == (h ^ (h >> 16))
    st, f = R.pcg_next(s)
    st2, f2 = R.pcg_next(st)
    assert np.all((f >= 0) & (f < 1)) and np.all((f2 >= 0) & (f2 < 1)) and not np.array_equal(f, f2)
    assert np.array_equal(st, ((s.astype(np.uint64) * 747796405 + 2891336453) & M).astype(np.uint32))


def test_camera_rays_through_an_identity_projection_point_at_the_far_plane():
    o, d = R.camera_rays(np.eye(4, dtype=f32).T.reshape(-1), (0, 0, 0), 8, 6, sample_index=3)
    assert d.shape == (48, 3) and np.allclose(np.linalg.norm(d, axis=1), 1, atol=1e-6)
    # ndc = (2u-1, -(2v-1), 1): the first pixel looks up-left, the last down-right, all towards +z (depth 1)
    assert d[0, 0] < 0 < d[0, 1] and d[-1, 0] > 0 > d[-1, 1] and np.all(d[:, 2] > 0)
    o2, d2 = R.camera_rays(np.eye(4, dtype=f32).reshape(-1), (0, 0, 0), 8, 6, sample_index=3, rect=(2, 1, 4, 3))
    full = d.reshape(6, 8, 3)[1:4, 2:6].reshape(-1, 3)
    assert np.array_equal(d2, full)  # a tile of the launch is the same rays


def test_tbn_is_orthonormal_and_ao_directions_stay_in_the_hemisphere():
    rng = np.random.default_rng(5)
    n = R._normalize(rng.normal(size=(256, 3)).astype(f32))
    n[0] = (0, 0, 1); n[1] = (0, 0, -1); n[2] = (1, 0, 0)
    ex = R.tbn_mul(n, np.tile(np.array([[1, 0, 0]], f32), (256, 1)))
    ey = R.tbn_mul(n, np.tile(np.array([[0, 1, 0]], f32), (256, 1)))
    ez = R.tbn_mul(n, np.tile(np.array([[0, 0, 1]], f32), (256, 1)))
    assert np.allclose(ez, n)
    for a, b in ((ex, ey), (ex, n), (ey, n)):
        assert np.max(np.abs(np.sum(a * b, axis=1))) < 1e-5
    for k in (0, 1, 17, 255):
        d = R.ao_directions(n, k)
        assert np.all(np.sum(d * n, axis=1) >= -1e-6) and np.allclose(np.linalg.norm(d, axis=1), 1, atol=1e-5)


def test_cosine_bounce_restatement_equals_the_scene_generator():
    # the product-side synthetic-input generator (scenes.py) and the oracle restatement are the same recipe
    idx = np.arange(0, 5000, 7, dtype=np.uint32)
    s = np.stack([R.van_der_corput(idx, S.SCRAMBLE_VDC), R.sobol(idx, S.SCRAMBLE_SOBOL)], -1)
    assert np.array_equal(s, S.sample_2d(idx))
    rng = np.random.default_rng(11)
    n = R._normalize(rng.normal(size=(idx.size, 3)).astype(f32))
    a, b = R.cosine_sample_hemisphere_in_dir(n, s), S.cosine_sample_hemisphere_in_dir(n, s)
    assert np.max(np.abs(a - b)) < 2e-6 and np.all(np.sum(a * n, axis=1) > -1e-6)


def test_geometric_normal_faces_the_ray_and_follows_the_instance_transform():
    pos, idx = S.torus_mesh(24, 16, 1.0, 0.35)  # no zero-area triangles (the UV sphere has them at its poles)
    m = S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(2, 3, 4))
    import oracle
    w2o = oracle.mat4_inverse_or_identity(m)
    prim = np.arange(40, 200, dtype=np.int64)
    tri = idx.reshape(-1, 3)[prim]
    centre = S.mat4_apply_point(m, pos[tri].mean(1).astype(f32))
    g = R.geometric_normals(pos, idx, prim, w2o, np.zeros(3, f32), centre)
    assert np.allclose(np.linalg.norm(g, axis=1), 1, atol=1e-5)
    assert np.all(np.sum((np.zeros(3, f32) - centre) * g, axis=1) >= 0)
    # against the world-space construction used by the scene generator (same direction up to rounding)
    d = R._normalize(centre)
    g2 = S.geometric_normals(pos, idx, prim, m, d)
    assert np.max(np.abs(g - g2)) < 1e-4


def test_offset_ray_hit_known_answers():
    """ray_util.rs:6-40: integer offset of int(256 n) ulps away from the origin side for |p| >= 1/32, n/65536 below"""
    p = np.array([[1.0, -2.0, 0.01], [100.0, 0.5, -0.03], [-1e-3, 1 / 32, -(1 / 32)]], f32)
    n = np.array([[0.0, 1.0, 0.0], [0.6, -0.8, 0.0], [1.0, 0.5, -0.25]], f32)
    out = R.offset_ray_hit(p, n)
    # row 0: x: of_i = 0 -> unchanged; y: p<0, of_i=256 -> bits - 256 (towards zero = along +n); z: |p|<1/32 -> p + n/65536 = p
    assert out[0, 0] == f32(1.0)
    assert out[0, 1].view(np.int32) == p[0, 1].view(np.int32) - 256 and out[0, 1] > p[0, 1]
    assert out[0, 2] == p[0, 2]
    # row 1: int(153.6) = 153 ulps up; int(-204.8) = -204 ulps (truncation, not floor)
    assert out[1, 0].view(np.int32) == p[1, 0].view(np.int32) + 153
    assert out[1, 1].view(np.int32) == p[1, 1].view(np.int32) - 204 and out[1, 1] < p[1, 1]
    assert out[1, 2] == (p[1, 2] + f32(1 / 65536) * n[1, 2]).astype(f32)
    # row 2: the 1/32 boundary is exclusive on the small side
    assert out[2, 0] == (p[2, 0] + f32(1 / 65536) * n[2, 0]).astype(f32)
    assert out[2, 1].view(np.int32) == p[2, 1].view(np.int32) + 128
    assert out[2, 2].view(np.int32) == p[2, 2].view(np.int32) + 64  # p<0: -of_i = +64 -> larger magnitude = along -z = along n
    # the offset always moves along the normal's sign, component by component
    rng = np.random.default_rng(5)
    P = rng.uniform(-50, 50, (4096, 3)).astype(f32)
    N = rng.normal(size=(4096, 3)).astype(f32); N /= np.linalg.norm(N, axis=1, keepdims=True)
    D = R.offset_ray_hit(P, N).astype(np.float64) - P
    assert np.all(D * N >= 0) and np.all(np.abs(D) <= np.abs(P) * 4e-5 + 2e-5)


def test_towards_point():
    d, dist = R.towards_point(np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]], f32), (4.0, 6.0, 3.0))
    assert np.array_equal(dist, np.array([5.0, np.sqrt(f32(61.0))], f32))
    assert np.array_equal(d[0], np.array([0.6, 0.8, 0.0], f32))
