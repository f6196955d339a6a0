"""Synthetic scene / ray generators for the BASELINE.json configs (inputs only — not on the hot path).

Everything is deterministic float32 numpy.  The mesh topology conventions follow the reference's
tessellator (vertex index = v + (V+1)*u, two triangles (a,c,b) and (b,c,d) per quad —
content/mesh/generator/src/builder/mod.rs:73-126); the surfaces are the reference's UV sphere
(content/mesh/generator/src/primitive.rs:23-35) and torus (builtin.rs:4-21) restated analytically.
Bit-identity with Rust's sin_cos is NOT required: the mesh is an input shared by the oracle and the
CUDA path.

Ray recipes:
  pinhole      shader/ray-tracing/.../geometry/naive/test.rs:259-264
  jittered     scene/rendering/gpu-ray-tracing/src/camera.rs:72-73 (uv + sample/size)
  cosine bounce  math/statistics/src/distribution_map.rs:10-58 with the (van_der_corput, sobol)
                 pair of math/statistics/src/sampling/sobol.rs:40-68 and FIXED scrambles
                 (the reference seeds them from rand::rng(), sobol.rs:10-18)
"""
from __future__ import annotations

import numpy as np

RAY_DTYPE = np.dtype([("ox", "f4"), ("oy", "f4"), ("oz", "f4"), ("tmin", "f4"),
                      ("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("tmax", "f4")])
INSTANCE_DTYPE = np.dtype([("transform", "f4", (16,)), ("instance_custom_index", "u4"), ("mask", "u4"),
                           ("sbt_offset", "u4"), ("flags", "u4"), ("blas_handle", "u4")])

SCRAMBLE_VDC = 0x9E3779B9
SCRAMBLE_SOBOL = 0x85EBCA6B

f32 = np.float32


# ----------------------------------------------------------------------------- matrices (column-major a1..d4)
def mat4_identity() -> np.ndarray:
    return np.eye(4, dtype=f32).T.reshape(16).copy()


def mat4_translate(x, y, z) -> np.ndarray:
    m = mat4_identity(); m[12], m[13], m[14] = f32(x), f32(y), f32(z); return m


def mat4_scale(x, y, z) -> np.ndarray:
    m = mat4_identity(); m[0], m[5], m[10] = f32(x), f32(y), f32(z); return m


def mat4_rotate_x(t) -> np.ndarray:
    s, c = f32(np.sin(f32(t))), f32(np.cos(f32(t)))
    m = mat4_identity(); m[5], m[6], m[9], m[10] = c, s, -s, c; return m


def mat4_rotate_y(t) -> np.ndarray:
    s, c = f32(np.sin(f32(t))), f32(np.cos(f32(t)))
    m = mat4_identity(); m[0], m[2], m[8], m[10] = c, -s, s, c; return m


def mat4_rotate_z(t) -> np.ndarray:
    s, c = f32(np.sin(f32(t))), f32(np.cos(f32(t)))
    m = mat4_identity(); m[0], m[1], m[4], m[5] = c, s, -s, c; return m


def mat4_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a * b in the reference's column-major convention, f32, term order of mat4.rs:176-204."""
    A = a.reshape(4, 4)  # A[col, row]
    B = b.reshape(4, 4)
    out = np.zeros((4, 4), f32)
    for col in range(4):
        for row in range(4):
            acc = f32(A[0, row] * B[col, 0])
            acc = f32(acc + f32(A[1, row] * B[col, 1]))
            acc = f32(acc + f32(A[2, row] * B[col, 2]))
            acc = f32(acc + f32(A[3, row] * B[col, 3]))
            out[col, row] = acc
    return out.reshape(16)


def mat4_apply_point(m: np.ndarray, p: np.ndarray) -> np.ndarray:
    """Mat4 * Vec3 (affine + /w) over an [n,3] array, f32."""
    M = m.reshape(4, 4)
    p = p.astype(f32)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    out = np.empty_like(p)
    w = (x * M[0, 3] + y * M[1, 3] + z * M[2, 3] + M[3, 3]).astype(f32)
    for r in range(3):
        out[:, r] = ((x * M[0, r] + y * M[1, r] + z * M[2, r] + M[3, r]).astype(f32) / w).astype(f32)
    return out


# ----------------------------------------------------------------------------- meshes
def _grid_indices(U: int, V: int) -> np.ndarray:
    u = np.arange(U, dtype=np.int64)[:, None]
    v = np.arange(V, dtype=np.int64)[None, :]
    a = v + (V + 1) * u
    b = (v + 1) + (V + 1) * u
    c = v + (V + 1) * (u + 1)
    d = (v + 1) + (V + 1) * (u + 1)
    tris = np.stack([a, c, b, b, c, d], axis=-1).reshape(-1, 3)
    return tris.astype(np.uint32)


def _grid_uv(U: int, V: int):
    u_step = f32(1.0) / f32(U)
    v_step = f32(1.0) / f32(V)
    u = (np.arange(U + 1, dtype=f32) * u_step).astype(f32)
    v = (np.arange(V + 1, dtype=f32) * v_step).astype(f32)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    return uu.reshape(-1), vv.reshape(-1)


def uv_sphere_mesh(U: int, V: int, radius: float = 1.0):
    """UV sphere: (cos(2*pi*u) sin(pi*v), cos(pi*v), sin(2*pi*u) sin(pi*v)) * radius.
    (U+1)*(V+1) vertices, 2*U*V triangles; pole triangles are degenerate exactly like the reference's."""
    uu, vv = _grid_uv(U, V)
    pi = f32(np.pi)
    su, cu = np.sin(uu * pi * f32(2.0)).astype(f32), np.cos(uu * pi * f32(2.0)).astype(f32)
    sv, cv = np.sin(vv * pi).astype(f32), np.cos(vv * pi).astype(f32)
    pos = np.stack([cu * sv, cv, su * sv], axis=-1).astype(f32) * f32(radius)
    return np.ascontiguousarray(pos.astype(f32)), _grid_indices(U, V)


def torus_mesh(U: int, V: int, radius: float = 1.0, tube_radius: float = 0.35):
    """Torus around the z axis: ((R + r cos p) cos t, (R + r cos p) sin t, r sin p), t = 2*pi*u, p = 2*pi*v.
    Outward-facing winding under the (a,c,b)/(b,c,d) convention."""
    uu, vv = _grid_uv(U, V)
    two_pi = f32(2.0) * f32(np.pi)
    st, ct = np.sin(uu * two_pi).astype(f32), np.cos(uu * two_pi).astype(f32)
    sp, cp = np.sin(vv * two_pi).astype(f32), np.cos(vv * two_pi).astype(f32)
    ring = (f32(radius) + f32(tube_radius) * cp).astype(f32)
    pos = np.stack([ring * ct, ring * st, f32(tube_radius) * sp], axis=-1).astype(f32)
    return np.ascontiguousarray(pos), _grid_indices(U, V)


CUBE_POSITION = np.array([
    0.5, 0.5, 0.5, -0.5, 0.5, 0.5, -0.5, -0.5, 0.5, 0.5, -0.5, 0.5,
    0.5, 0.5, 0.5, 0.5, -0.5, 0.5, 0.5, -0.5, -0.5, 0.5, 0.5, -0.5,
    0.5, 0.5, 0.5, 0.5, 0.5, -0.5, -0.5, 0.5, -0.5, -0.5, 0.5, 0.5,
    -0.5, 0.5, 0.5, -0.5, 0.5, -0.5, -0.5, -0.5, -0.5, -0.5, -0.5, 0.5,
    -0.5, -0.5, -0.5, 0.5, -0.5, -0.5, 0.5, -0.5, 0.5, -0.5, -0.5, 0.5,
    0.5, -0.5, -0.5, -0.5, -0.5, -0.5, -0.5, 0.5, -0.5, 0.5, 0.5, -0.5], dtype=f32).reshape(-1, 3)
CUBE_INDEX = np.array([0, 1, 2, 2, 3, 0, 4, 5, 6, 6, 7, 4, 8, 9, 10, 10, 11, 8, 12, 13, 14, 14, 15, 12,
                       16, 17, 18, 18, 19, 16, 20, 21, 22, 22, 23, 20], dtype=np.uint32)
"""The reference fixture's literal cube (naive/test.rs:80-96)."""


def make_instance(transform, blas_handle, custom_index=0, mask=0xFFFFFFFF, flags=0, sbt_offset=0) -> np.ndarray:
    inst = np.zeros(1, INSTANCE_DTYPE)
    inst["transform"][0] = np.asarray(transform, f32).reshape(16)
    inst["instance_custom_index"] = custom_index
    inst["mask"] = mask
    inst["sbt_offset"] = sbt_offset
    inst["flags"] = flags
    inst["blas_handle"] = blas_handle
    return inst


def instance_grid(nx: int, ny: int, blas_handle: int, spacing: float = 3.5, z: float = -200.0) -> np.ndarray:
    """BASELINE config 4: transform_i = translate(s*(ix-cx), s*(iy-cy), z) * rotate_y(0.37 i) * scale(1 + 0.25 (i%3))."""
    out = np.zeros(nx * ny, INSTANCE_DTYPE)
    cx, cy = (nx - 1) / 2.0, (ny - 1) / 2.0
    for iy in range(ny):
        for ix in range(nx):
            i = iy * nx + ix
            s = 1.0 + 0.25 * (i % 3)
            m = mat4_mul(mat4_mul(mat4_translate(spacing * (ix - cx), spacing * (iy - cy), z), mat4_rotate_y(0.37 * i)),
                         mat4_scale(s, s, s))
            out[i] = make_instance(m, blas_handle, custom_index=i)[0]
    return out


# ----------------------------------------------------------------------------- rays
def _normalize(v: np.ndarray) -> np.ndarray:
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    mag = ((x * x).astype(f32) + (y * y).astype(f32)).astype(f32)
    mag = (mag + (z * z).astype(f32)).astype(f32)
    inv = (f32(1.0) / np.sqrt(mag, dtype=f32)).astype(f32)
    inv = np.where(mag > 0, inv, f32(1.0)).astype(f32)
    return (v * inv[..., None]).astype(f32)


def make_rays(origin: np.ndarray, direction: np.ndarray, tmin: float, tmax: float) -> np.ndarray:
    n = direction.shape[0]
    rays = np.zeros(n, RAY_DTYPE)
    o = np.broadcast_to(np.asarray(origin, f32), (n, 3))
    rays["ox"], rays["oy"], rays["oz"] = o[:, 0], o[:, 1], o[:, 2]
    rays["dx"], rays["dy"], rays["dz"] = direction[:, 0], direction[:, 1], direction[:, 2]
    rays["tmin"], rays["tmax"] = f32(tmin), f32(tmax)
    return rays


def pinhole_rays(W: int, H: int, tmin: float = 0.0, tmax: float = 100.0, origin=(0.0, 0.0, 0.0),
                 aspect_correct: bool = False, jitter: np.ndarray | None = None, rows: tuple[int, int] | None = None) -> np.ndarray:
    """fov-90 pinhole grid, row-major (j outer, i inner): x = (i+.5)/W*2-1, y = 1-(j+.5)/H*2, target=(x,y,-1).
    `jitter` ([n,2] in [0,1)) replaces the .5 pixel-centre offset; `rows` = (j0, j1) restricts to a row band."""
    j0, j1 = rows if rows is not None else (0, H)
    i = np.arange(W, dtype=f32)[None, :].repeat(j1 - j0, axis=0).reshape(-1)
    j = np.arange(j0, j1, dtype=f32)[:, None].repeat(W, axis=1).reshape(-1)
    if jitter is None:
        jx = jy = f32(0.5)
    else:
        jx, jy = jitter[:, 0].astype(f32), jitter[:, 1].astype(f32)
    x = ((i + jx).astype(f32) / f32(W) * f32(2.0) - f32(1.0)).astype(f32)
    y = (f32(1.0) - (j + jy).astype(f32) / f32(H) * f32(2.0)).astype(f32)
    if aspect_correct:
        x = (x * f32(W / H)).astype(f32)
    o = np.asarray(origin, f32)
    target = np.stack([x, y, np.full_like(x, f32(-1.0))], axis=-1)
    d = _normalize((target - o[None, :]).astype(f32))
    return make_rays(o, d, tmin, tmax)


def van_der_corput(n: np.ndarray, scramble: int) -> np.ndarray:
    n = n.astype(np.uint32)
    n = (n >> np.uint32(16)) | (n << np.uint32(16))
    n = ((n & np.uint32(0x00ff00ff)) << np.uint32(8)) | ((n & np.uint32(0xff00ff00)) >> np.uint32(8))
    n = ((n & np.uint32(0x0f0f0f0f)) << np.uint32(4)) | ((n & np.uint32(0xf0f0f0f0)) >> np.uint32(4))
    n = ((n & np.uint32(0x33333333)) << np.uint32(2)) | ((n & np.uint32(0xcccccccc)) >> np.uint32(2))
    n = ((n & np.uint32(0x55555555)) << np.uint32(1)) | ((n & np.uint32(0xaaaaaaaa)) >> np.uint32(1))
    n = n ^ np.uint32(scramble)
    v = ((n >> np.uint32(8)) & np.uint32(0xffffff)).astype(f32) / f32(1 << 24)
    return np.minimum(v, f32(1.0) - np.finfo(f32).eps).astype(f32)


def sobol(n: np.ndarray, scramble: int) -> np.ndarray:
    n = n.astype(np.uint32).copy()
    s = np.full(n.shape, scramble, np.uint32)
    i = np.uint32(1 << 31)
    for _ in range(32):
        s = np.where((n & np.uint32(1)) != 0, s ^ i, s)
        n = n >> np.uint32(1)
        i = i ^ (i >> np.uint32(1))
    v = ((s >> np.uint32(8)) & np.uint32(0xffffff)).astype(f32) / f32(1 << 24)
    return np.minimum(v, f32(1.0) - np.finfo(f32).eps).astype(f32)


def sample_2d(n: np.ndarray) -> np.ndarray:
    """SobolSamplingGenerator::gen_2d with fixed scrambles."""
    return np.stack([van_der_corput(n, SCRAMBLE_VDC), sobol(n, SCRAMBLE_SOBOL)], axis=-1)


def cosine_sample_hemisphere_in_dir(direction: np.ndarray, sample: np.ndarray) -> np.ndarray:
    """distribution_map.rs:10-58 vectorised in f32."""
    d = direction.astype(f32)
    ux = (sample[:, 0] * f32(2.0) - f32(1.0)).astype(f32)
    uy = (sample[:, 1] * f32(2.0) - f32(1.0)).astype(f32)
    pi4, pi2 = f32(np.pi / 4), f32(np.pi / 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.abs(ux) > np.abs(uy)
        r = np.where(a, ux, uy).astype(f32)
        theta = np.where(a, pi4 * (uy / ux), pi2 - pi4 * (ux / uy)).astype(f32)
    zero = (ux == 0) & (uy == 0)
    dx = np.where(zero, f32(0), np.cos(theta).astype(f32) * r).astype(f32)
    dy = np.where(zero, f32(0), np.sin(theta).astype(f32) * r).astype(f32)
    z = np.sqrt(np.maximum(f32(0), f32(1.0) - dx * dx - dy * dy).astype(f32)).astype(f32)
    up_y = np.broadcast_to(np.array([0, 1, 0], f32), d.shape)
    left = _normalize(np.cross(up_y, d).astype(f32))
    up = np.cross(left, d).astype(f32)
    xy_r = np.sqrt((dx * dx + dy * dy).astype(f32)).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        cos_phi = (dx / xy_r).astype(f32)
        sin_phi = (dy / xy_r).astype(f32)
    out = (left * (xy_r * cos_phi)[:, None] + up * (xy_r * sin_phi)[:, None] + d * z[:, None]).astype(f32)
    out = _normalize(out)
    out = np.where((xy_r == 0)[:, None], d, out).astype(f32)
    return np.ascontiguousarray(out)


def bounce_rays(primary: np.ndarray, hits: np.ndarray, normals: np.ndarray, tmin: float = 0.01, tmax: float = 100.0):
    """One cosine-weighted bounce per primary hit (BASELINE config 3).  `normals` = geometric normal per ray
    (already flipped to face the incoming ray).  Returns (rays, index_of_primary)."""
    hit = hits["instance_id"] != 0xFFFFFFFF
    idx = np.nonzero(hit)[0]
    o = np.stack([primary["ox"], primary["oy"], primary["oz"]], -1)[idx]
    d = np.stack([primary["dx"], primary["dy"], primary["dz"]], -1)[idx]
    t = hits["t"][idx].astype(f32)
    p = (o + d * t[:, None]).astype(f32)
    s = sample_2d(idx.astype(np.uint32))
    nd = cosine_sample_hemisphere_in_dir(normals[idx].astype(f32), s)
    rays = np.zeros(idx.size, RAY_DTYPE)
    rays["ox"], rays["oy"], rays["oz"] = p[:, 0], p[:, 1], p[:, 2]
    rays["dx"], rays["dy"], rays["dz"] = nd[:, 0], nd[:, 1], nd[:, 2]
    rays["tmin"], rays["tmax"] = f32(tmin), f32(tmax)
    return rays, idx


def geometric_normals(positions: np.ndarray, indices: np.ndarray, prim: np.ndarray, transform: np.ndarray | None,
                      ray_dir: np.ndarray) -> np.ndarray:
    """World-space geometric normal of triangle `prim`, flipped to face -ray_dir."""
    tri = indices.reshape(-1, 3)[prim]
    P = positions if transform is None else mat4_apply_point(transform, positions)
    v0, v1, v2 = P[tri[:, 0]], P[tri[:, 1]], P[tri[:, 2]]
    n = _normalize(np.cross((v1 - v0).astype(f32), (v2 - v0).astype(f32)).astype(f32))
    flip = np.sum(n * ray_dir, axis=-1) > 0
    n[flip] = -n[flip]
    return n.astype(f32)
