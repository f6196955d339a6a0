"""oracle/raygen.py — numpy restatement of the reference's ray-generation and bounce recipes (TEST INFRASTRUCTURE ONLY:
imported by tests/, bench.py's checker legs and __graft_entry__.smoke(), never by the product).

Every function is plain f32 arithmetic in the order the reference writes it; each cites the lines it follows.  The
reference runs these as WGSL on the GPU (no golden values exist upstream: "parity unpinned"), so the device kernels of
rendiation_b200/csrc/raygen.cu are compared to this file within a few ulp (sin/cos differ between libms), and the
traversal parity is then checked on exactly the rays the device produced.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
u32 = np.uint32


def _normalize(v):
    """InnerProductSpace::normalize (math/algebra/src/vec/dimension.rs:98-105): v * (1/sqrt(len2)) if len2 > 0 else v"""
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    mag = ((x * x).astype(f32) + (y * y).astype(f32)).astype(f32)
    mag = (mag + (z * z).astype(f32)).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (f32(1.0) / np.sqrt(mag, dtype=f32)).astype(f32)
    inv = np.where(mag > 0, inv, f32(1.0)).astype(f32)
    return (v * inv[..., None]).astype(f32)


def _dot(a, b):
    return ((a[..., 0] * b[..., 0]).astype(f32) + (a[..., 1] * b[..., 1]).astype(f32) + (a[..., 2] * b[..., 2]).astype(f32)).astype(f32)


def _cross(a, b):
    return np.stack([(a[..., 1] * b[..., 2]).astype(f32) - (a[..., 2] * b[..., 1]).astype(f32),
                     (a[..., 2] * b[..., 0]).astype(f32) - (a[..., 0] * b[..., 2]).astype(f32),
                     (a[..., 0] * b[..., 1]).astype(f32) - (a[..., 1] * b[..., 0]).astype(f32)], -1).astype(f32)


# ---------------------------------------------------------------------------------------------- samplers
def xxhash32(px, py, pz):
    """scene/rendering/gpu-ray-tracing/src/sampler.rs:27-42"""
    px, py, pz = (np.asarray(a, np.uint64) for a in (px, py, pz))
    M = np.uint64(0xFFFFFFFF)
    p0, p1, p2, p3 = (np.uint64(v) for v in (2246822519, 3266489917, 668265263, 374761393))

    def rotl17(h):
        return ((h << np.uint64(17)) | (h >> np.uint64(15))) & M

    h = (pz + p3 + (px * p1 & M)) & M
    h = (p2 * rotl17(h)) & M
    h = (h + (py * p1 & M)) & M
    h = (p2 * rotl17(h)) & M
    h = (p0 * (h ^ (h >> np.uint64(15)))) & M
    h = (p1 * (h ^ (h >> np.uint64(13)))) & M
    return (h ^ (h >> np.uint64(16))).astype(u32)


def pcg_next(state):
    """pcg + PCGRandomSampler::next (sampler.rs:11-16, 67-71): returns (new_state, float in [0,1))"""
    s = np.asarray(state, np.uint64)
    M = np.uint64(0xFFFFFFFF)
    prev = (s * np.uint64(747796405) + np.uint64(2891336453)) & M
    word = (((prev >> ((prev >> np.uint64(28)) + np.uint64(4))) ^ prev) * np.uint64(277803737)) & M
    r = ((word >> np.uint64(22)) ^ word) & M
    bits = (np.uint64(0x3f800000) | (r >> np.uint64(9))).astype(u32)
    return prev.astype(u32), (bits.view(f32) - f32(1.0)).astype(f32)


def radical_inverse_vdc(bits):
    """shader/library/src/sampling.rs:43-52"""
    b = np.asarray(bits, u32)
    b = (b << u32(16)) | (b >> u32(16))
    b = ((b & u32(0x55555555)) << u32(1)) | ((b & u32(0xAAAAAAAA)) >> u32(1))
    b = ((b & u32(0x33333333)) << u32(2)) | ((b & u32(0xCCCCCCCC)) >> u32(2))
    b = ((b & u32(0x0F0F0F0F)) << u32(4)) | ((b & u32(0xF0F0F0F0)) >> u32(4))
    b = ((b & u32(0x00FF00FF)) << u32(8)) | ((b & u32(0xFF00FF00)) >> u32(8))
    return (b.astype(f32) * f32(2.3283064e-10)).astype(f32)


def hammersley_2d(index, total):
    """shader/library/src/sampling.rs:33-41"""
    index = np.asarray(index, u32)
    return np.stack([(index.astype(f32) / f32(total)).astype(f32), radical_inverse_vdc(index)], -1)


# ---------------------------------------------------------------------------------------------- primary rays
def camera_rays(view_projection_inv, world_position, width, height, sample_index=0, rect=None, ndc_depth=1.0):
    """DefaultRtxCameraInvocation::generate_ray (scene/rendering/gpu-ray-tracing/src/camera.rs:66-98) with
    PCGRandomSampler::from_ray_ctx_and_sample_index (sampler.rs:49-57) and shader_uv_space_to_render_space
    (shader/library/src/lib.rs:18-28).  Returns (origin[3], directions[n,3]) for the rectangle, row-major."""
    x0, y0, w, h = rect if rect is not None else (0, 0, width, height)
    px = np.tile(np.arange(x0, x0 + w, dtype=u32), h)
    py = np.repeat(np.arange(y0, y0 + h, dtype=u32), w)
    state = xxhash32(px, py, np.full_like(px, sample_index))
    state, s0 = pcg_next(state)
    state, s1 = pcg_next(state)
    fw, fh = f32(width), f32(height)
    u = ((px.astype(f32) / fw).astype(f32) + (s0 / fw).astype(f32)).astype(f32)
    v = ((py.astype(f32) / fh).astype(f32) + (s1 / fh).astype(f32)).astype(f32)
    nx = ((u * f32(2.0)).astype(f32) - f32(1.0)).astype(f32)
    ny = (((v * f32(2.0)).astype(f32) - f32(1.0)).astype(f32) * f32(-1.0)).astype(f32)
    nz = np.full_like(nx, f32(ndc_depth))
    nw = np.ones_like(nx)
    m = np.asarray(view_projection_inv, f32).reshape(-1)  # column-major a1..d4 (mat4.rs:9-14, Mat4*Vec4 :161-168)

    def row(k):
        acc = (nx * m[k]).astype(f32)
        acc = (acc + (ny * m[4 + k]).astype(f32)).astype(f32)
        acc = (acc + (nz * m[8 + k]).astype(f32)).astype(f32)
        return (acc + (nw * m[12 + k]).astype(f32)).astype(f32)

    x, y, z, wv = row(0), row(1), row(2), row(3)
    target = np.stack([(x / wv).astype(f32), (y / wv).astype(f32), (z / wv).astype(f32)], -1)
    o = np.asarray(world_position, f32)
    return o, _normalize((target - o[None, :]).astype(f32))


# ---------------------------------------------------------------------------------------------- closest hit -> bounce
def geometric_normals(positions, indices, prim, world_to_object, ray_origin, hit_position):
    """BindlessMeshRtxAccessInvocation::get_world_normal_impl, geometric part
    (scene/rendering/gpu-ray-tracing/src/bindless_mesh_bridge.rs:95-114): normalize(normal_mat * (pa-pb) x (pa-pc)) with
    normal_mat = transpose(mat3(world_to_object)), negated when dot(origin - hit, n) < 0."""
    tri = np.asarray(indices).reshape(-1, 3)[prim]
    P = np.asarray(positions, f32)
    pa, pb, pc = P[tri[:, 0]], P[tri[:, 1]], P[tri[:, 2]]
    c = _cross((pa - pb).astype(f32), (pa - pc).astype(f32))
    wi = np.asarray(world_to_object, f32).reshape(-1)  # column-major

    def comp(k):  # row k of the transpose = column k of world_to_object
        acc = (wi[4 * k] * c[:, 0]).astype(f32)
        acc = (acc + (wi[4 * k + 1] * c[:, 1]).astype(f32)).astype(f32)
        return (acc + (wi[4 * k + 2] * c[:, 2]).astype(f32)).astype(f32)

    g = _normalize(np.stack([comp(0), comp(1), comp(2)], -1))
    flip = _dot((np.asarray(ray_origin, f32) - hit_position).astype(f32), g) < 0
    g[flip] = -g[flip]
    return g


def offset_ray_hit(position, normal):
    """scene/rendering/gpu-ray-tracing/src/ray_util.rs:6-40 (Ray Tracing Gems ch. 6): per component, the position moved by
    int(256 * n) units in the last place along the normal, or by n / 65536 when |p| < 1/32"""
    p = np.asarray(position, f32)
    n = np.asarray(normal, f32)
    of_i = np.trunc((n * f32(256.0)).astype(f32)).astype(np.int32)  # into_i32
    step = np.where(p < 0, -of_i, of_i).astype(np.int32)
    p_i = (p.view(np.int32) + step).astype(np.int32).view(f32)
    near = (p + (f32(1.0 / 65536.0) * n).astype(f32)).astype(f32)
    return np.where(np.abs(p) < f32(1.0 / 32.0), near, p_i).astype(f32)


def towards_point(position, target):
    """PointLight::importance_sampling_light_impl (feature/path_tracing/lighting_bridge.rs:86-94): (direction, distance)"""
    to_light = (np.asarray(target, f32)[None, :] - np.asarray(position, f32)).astype(f32)
    acc = (to_light[:, 0] * to_light[:, 0]).astype(f32)
    acc = (acc + (to_light[:, 1] * to_light[:, 1]).astype(f32)).astype(f32)
    acc = (acc + (to_light[:, 2] * to_light[:, 2]).astype(f32)).astype(f32)
    distance = np.sqrt(acc, dtype=f32)
    return (to_light / distance[:, None]).astype(f32), distance


def sample_hemisphere_cos(uv):
    """shader/library/src/sampling.rs:55-62"""
    phi = (f32(2.0 * np.pi) * uv[..., 1]).astype(f32)
    cos_theta = np.sqrt((f32(1.0) - uv[..., 0]).astype(f32), dtype=f32)
    sin_theta = np.sqrt((f32(1.0) - (cos_theta * cos_theta).astype(f32)).astype(f32), dtype=f32)
    return np.stack([(np.cos(phi).astype(f32) * sin_theta).astype(f32), (np.sin(phi).astype(f32) * sin_theta).astype(f32), cos_theta], -1)


def tbn_mul(normal, local):
    """tbn(normal) * local (shader/library/src/sampling.rs:66-83, Pixar orthonormal basis; Mat3*Vec3 mat3.rs:103-109)"""
    n = np.asarray(normal, f32)
    sign = np.where(n[:, 2] < 0, f32(-1.0), f32(1.0)).astype(f32)
    a = (f32(-1.0) / (sign + n[:, 2]).astype(f32)).astype(f32)
    b = ((n[:, 0] * n[:, 1]).astype(f32) * a).astype(f32)
    tangent = _normalize(np.stack([(f32(1.0) + (((sign * n[:, 0]).astype(f32) * n[:, 0]).astype(f32) * a).astype(f32)).astype(f32),
                                   (sign * b).astype(f32), ((-sign) * n[:, 0]).astype(f32)], -1))
    bi = _normalize(np.stack([b, (sign + ((n[:, 1] * n[:, 1]).astype(f32) * a).astype(f32)).astype(f32), -n[:, 1]], -1))
    out = []
    for k in range(3):
        acc = (tangent[:, k] * local[:, 0]).astype(f32)
        acc = (acc + (bi[:, k] * local[:, 1]).astype(f32)).astype(f32)
        out.append((acc + (n[:, k] * local[:, 2]).astype(f32)).astype(f32))
    return np.stack(out, -1)


def ao_directions(normals, sample_index, max_sample=256):
    """the AO secondary ray direction (scene/rendering/gpu-ray-tracing/src/feature/ao.rs:258-265)"""
    uv = np.broadcast_to(hammersley_2d(np.array([sample_index], u32), max_sample), (normals.shape[0], 2))
    return tbn_mul(normals, sample_hemisphere_cos(uv))


# ---------------------------------------------------------------------------------------------- SURVEY §8d config-3 bounce
def van_der_corput(n, scramble):
    """SobolSamplingGenerator, dimension 0 (math/statistics/src/sampling/sobol.rs:40-54) with a fixed scramble"""
    n = np.asarray(n, u32)
    n = (n >> u32(16)) | (n << u32(16))
    n = ((n & u32(0x00ff00ff)) << u32(8)) | ((n & u32(0xff00ff00)) >> u32(8))
    n = ((n & u32(0x0f0f0f0f)) << u32(4)) | ((n & u32(0xf0f0f0f0)) >> u32(4))
    n = ((n & u32(0x33333333)) << u32(2)) | ((n & u32(0xcccccccc)) >> u32(2))
    n = ((n & u32(0x55555555)) << u32(1)) | ((n & u32(0xaaaaaaaa)) >> u32(1))
    n = n ^ u32(scramble)
    v = ((n >> u32(8)) & u32(0xffffff)).astype(f32) / f32(1 << 24)
    return np.minimum(v, f32(1.0) - np.finfo(f32).eps).astype(f32)


def sobol(n, scramble):
    """SobolSamplingGenerator, dimension 1 (sobol.rs:56-68) with a fixed scramble"""
    n = np.asarray(n, u32).copy()
    s = np.full(n.shape, scramble, u32)
    i = u32(1 << 31)
    for _ in range(32):
        s = np.where((n & u32(1)) != 0, s ^ i, s)
        n = n >> u32(1)
        i = i ^ (i >> u32(1))
    v = ((s >> u32(8)) & u32(0xffffff)).astype(f32) / f32(1 << 24)
    return np.minimum(v, f32(1.0) - np.finfo(f32).eps).astype(f32)


def cosine_sample_hemisphere_in_dir(direction, sample):
    """math/statistics/src/distribution_map.rs:10-58 (concentric disk lifted to the hemisphere around `direction`)"""
    d = np.asarray(direction, f32)
    ux = ((sample[:, 0] * f32(2.0)).astype(f32) - f32(1.0)).astype(f32)
    uy = ((sample[:, 1] * f32(2.0)).astype(f32) - f32(1.0)).astype(f32)
    pi4, pi2 = f32(np.pi / 4), f32(np.pi / 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.abs(ux) > np.abs(uy)
        r = np.where(a, ux, uy).astype(f32)
        theta = np.where(a, (pi4 * (uy / ux).astype(f32)).astype(f32), (pi2 - (pi4 * (ux / uy).astype(f32)).astype(f32)).astype(f32)).astype(f32)
    zero = (ux == 0) & (uy == 0)
    dx = np.where(zero, f32(0), (np.cos(theta).astype(f32) * r).astype(f32)).astype(f32)
    dy = np.where(zero, f32(0), (np.sin(theta).astype(f32) * r).astype(f32)).astype(f32)
    z = np.sqrt(np.maximum(f32(0), ((f32(1.0) - (dx * dx).astype(f32)).astype(f32) - (dy * dy).astype(f32)).astype(f32)), dtype=f32)
    up_y = np.broadcast_to(np.array([0, 1, 0], f32), d.shape)
    left = _normalize(_cross(up_y, d))
    up = _cross(left, d)
    xy_r = np.sqrt(((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32), dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        cos_phi = (dx / xy_r).astype(f32)
        sin_phi = (dy / xy_r).astype(f32)
    out = ((left * (xy_r * cos_phi).astype(f32)[:, None]).astype(f32) + (up * (xy_r * sin_phi).astype(f32)[:, None]).astype(f32)).astype(f32)
    out = _normalize((out + (d * z[:, None]).astype(f32)).astype(f32))
    return np.ascontiguousarray(np.where((xy_r == 0)[:, None], d, out).astype(f32))
