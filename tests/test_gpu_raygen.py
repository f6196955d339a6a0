"""GPU tests of the device-side ray generation and closest-hit -> bounce step (rendiation_b200/csrc/raygen.cu)
against the numpy restatement of the reference recipes (oracle/raygen.py), and of the whole device-resident wavefront
primary -> compact -> bounce -> trace, whose hits are compared bit for bit with the oracle traversal of the very rays the
device generated."""
import numpy as np
import pytest

import oracle
from oracle import raygen as R
from rendiation_b200 import api, scenes as S

import helpers

pytestmark = pytest.mark.gpu
f32 = np.float32
DIR_TOL = 3e-6  # sinf/cosf of CUDA vs numpy: a few ulp on unit vectors


def _dev_rays(n):
    import torch
    return torch.zeros((n, 32), dtype=torch.uint8, device="cuda")


def _np_rays(t):
    return t.cpu().numpy().view(S.RAY_DTYPE).reshape(-1)


def _dirs(r):
    return np.stack([r["dx"], r["dy"], r["dz"]], -1)


def _orig(r):
    return np.stack([r["ox"], r["oy"], r["oz"]], -1)


def test_pinhole_rays_equal_the_host_recipe_bit_for_bit():
    import torch
    sp, _ = helpers.sphere_c1(seg=16)
    st = torch.cuda.current_stream().cuda_stream
    for (W, H, aspect, jit, rect) in [(64, 48, False, None, None), (1920, 1080, True, (0.25, 0.75), None),
                                      (200, 120, True, (0.5, 0.5), (37, 11, 101, 57)), (1, 1, False, None, None)]:
        x0, y0, w, h = rect if rect else (0, 0, W, H)
        d = _dev_rays(w * h)
        n = sp.p.gen_pinhole_rays_device(d.data_ptr(), W, H, rect=rect, tmin=0.01, tmax=100.0, aspect=float(f32(W / H)) if aspect else 1.0,
                                         jitter=jit or (0.5, 0.5), stream=st)
        torch.cuda.synchronize()
        jitter = None if jit is None else np.tile(np.array([jit], f32), (W * H, 1))
        want = S.pinhole_rays(W, H, 0.01, 100.0, aspect_correct=aspect, jitter=jitter).reshape(H, W)[y0:y0 + h, x0:x0 + w].reshape(-1)
        assert n == w * h and _np_rays(d).tobytes() == want.tobytes(), (W, H, rect)
    with pytest.raises(api.RdnError):
        sp.p.gen_pinhole_rays_device(_dev_rays(4).data_ptr(), 4, 4, rect=(2, 2, 3, 1))
    # many rectangles (tiles x samples) in one launch == the same rectangles one by one
    W, H = 300, 170
    rects = [(0, 0, 128, 128), (128, 0, 128, 128), (256, 0, 44, 128), (0, 128, 128, 42), (256, 128, 44, 42)] * 2
    jits = [(0.5, 0.5)] * 5 + [(0.125, 0.875)] * 5
    total = sum(w * h for (_, _, w, h) in rects)
    d = _dev_rays(total)
    assert sp.p.gen_pinhole_rays_batch_device(d.data_ptr(), W, H, rects, jits, tmin=0.01, tmax=100.0, aspect=float(f32(W / H)), stream=st) == total
    torch.cuda.synchronize()
    got, off = _np_rays(d), 0
    for (x0, y0, w, h), jit in zip(rects, jits):
        full = S.pinhole_rays(W, H, 0.01, 100.0, aspect_correct=True, jitter=np.tile(np.array([jit], f32), (W * H, 1)))
        want = full.reshape(H, W)[y0:y0 + h, x0:x0 + w].reshape(-1)
        assert got[off:off + w * h].tobytes() == want.tobytes(), (x0, y0)
        off += w * h


def test_camera_rays_equal_the_restated_ray_gen_shader():
    import torch
    sp, _ = helpers.sphere_c1(seg=16)
    st = torch.cuda.current_stream().cuda_stream
    # perspective * view, inverted with the reference's cofactor inverse (restated in the oracle)
    fov, asp, zn, zf = 1.0, 16 / 9, 0.1, 1000.0
    t = np.tan(fov / 2)
    proj = np.zeros((4, 4), f32)
    proj[0, 0] = 1 / (asp * t); proj[1, 1] = 1 / t; proj[2, 2] = zf / (zn - zf); proj[2, 3] = zn * zf / (zn - zf); proj[3, 2] = -1
    view = S.mat4_mul(S.mat4_rotate_y(0.3), S.mat4_translate(0.5, -0.25, 3.0))
    vp = S.mat4_mul(np.ascontiguousarray(proj.T).reshape(-1), view)  # column-major
    vp_inv = oracle.mat4_inverse_or_identity(vp)
    eye = (1.0, 2.0, 3.0)
    for sample, rect in [(0, None), (7, (5, 3, 40, 20)), (255, None)]:
        W, H = 96, 54
        x0, y0, w, h = rect if rect else (0, 0, W, H)
        d = _dev_rays(w * h)
        sp.p.gen_camera_rays_device(d.data_ptr(), vp_inv, eye, W, H, sample_index=sample, rect=rect, tmin=0.0, tmax=1e30, stream=st)
        torch.cuda.synchronize()
        got = _np_rays(d)
        o, dirs = R.camera_rays(vp_inv, eye, W, H, sample_index=sample, rect=rect)
        assert np.array_equal(_orig(got), np.tile(o, (w * h, 1)))
        assert np.array_equal(_dirs(got), dirs), float(np.max(np.abs(_dirs(got) - dirs)))
        assert np.all(got["tmin"] == 0.0) and np.all(got["tmax"] == f32(1e30))


def _primary_and_hits(sp, W, H):
    import torch
    st = torch.cuda.current_stream().cuda_stream
    d_rays = _dev_rays(W * H)
    d_hits = torch.zeros_like(d_rays)
    sp.p.gen_pinhole_rays_device(d_rays.data_ptr(), W, H, tmin=0.01, tmax=100.0, stream=st)
    sp.p.trace_closest_device(d_rays.data_ptr(), W * H, d_hits.data_ptr(), ray_flags=helpers.CULL_BACK, grid_width=W, stream=st)
    torch.cuda.synchronize()
    return d_rays, d_hits


@pytest.mark.parametrize("mode", [0, 1])
def test_bounce_rays_follow_the_reference_recipe(mode):
    import torch
    sp, (pos, idx, m) = helpers.torus_scene(96)
    W, H = 192, 160
    d_rays, d_hits = _primary_and_hits(sp, W, H)
    n = W * H
    st = torch.cuda.current_stream().cuda_stream
    d_out = _dev_rays(n)
    d_src = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    sp.p.gen_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_out.data_ptr(), d_src.data_ptr(), d_n.data_ptr(), mode=mode,
                                index_base=1000, sample_index=37, max_sample=256, tmin=0.01, tmax=100.0, stream=st)
    torch.cuda.synchronize()
    rays = _np_rays(d_rays)
    hits = d_hits.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
    src_want = np.nonzero(hits["instance_id"] != api.INVALID_ID)[0]
    k = int(d_n.item())
    assert k == src_want.size and k > 1000
    src = d_src.cpu().numpy().view(np.uint32)
    assert np.array_equal(src[:k], src_want) and np.all(src[k:] == 0)  # stable compaction, zero tail (the reference's contract)
    got = _np_rays(d_out)[:k]
    o, d, t = _orig(rays)[src_want], _dirs(rays)[src_want], hits["t"][src_want]
    p = (o + (d * t[:, None]).astype(f32)).astype(f32)
    assert np.array_equal(_orig(got), p)  # hit_world_position = origin + direction * t
    w2o = oracle.mat4_inverse_or_identity(m)
    g = R.geometric_normals(pos, idx, hits["primitive_id"][src_want].astype(np.int64), w2o, o, p)
    if mode == 0:
        s = np.stack([R.van_der_corput((src_want + 1000).astype(np.uint32), S.SCRAMBLE_VDC),
                      R.sobol((src_want + 1000).astype(np.uint32), S.SCRAMBLE_SOBOL)], -1)
        want = R.cosine_sample_hemisphere_in_dir(g, s)
    else:
        want = R.ao_directions(g, 37, 256)
    err = np.abs(_dirs(got) - want)
    # mode 0 lifts the disk sample with z = sqrt(1 - dx^2 - dy^2): near the rim the cancellation amplifies the few-ulp
    # sinf/cosf differences by 1/z, so the bulk is held to DIR_TOL and the worst (grazing) sample to 10x that
    assert float(np.quantile(err.max(axis=1), 0.999)) < DIR_TOL and float(err.max()) < 10 * DIR_TOL, (float(err.max()),)
    assert np.all(np.sum(_dirs(got) * g, axis=1) > -1e-5)  # bounce directions leave the surface on the ray's side
    assert np.all(got["tmin"] == f32(0.01)) and np.all(got["tmax"] == f32(100.0))


def test_shadow_test_rays_with_offset_origin_are_bit_exact():
    """mode 2 + RDN_BOUNCE_OFFSET_ORIGIN: the path tracer's shadow-test ray (ray_hit.rs:20-45) — integer origin offset, division by the
    distance — is plain f32 / integer arithmetic, so the device rays equal the numpy restatement bit for bit; traced with
    ACCEPT_FIRST_HIT_AND_END_SEARCH they equal the oracle's first-hit walk"""
    import torch
    sp, (pos, idx, m) = helpers.torus_scene(96)
    W, H = 192, 160
    d_rays, d_hits = _primary_and_hits(sp, W, H)
    n = W * H
    st = torch.cuda.current_stream().cuda_stream
    d_out = _dev_rays(n)
    d_src = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    light = (3.0, 8.0, -4.0)
    sp.p.gen_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_out.data_ptr(), d_src.data_ptr(), d_n.data_ptr(), mode=2,
                                tmin=float(np.finfo(f32).eps), target=light, offset_origin=True, stream=st)
    k = int(d_n.item())
    rays = _np_rays(d_rays)
    hits = d_hits.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
    src = np.nonzero(hits["instance_id"] != api.INVALID_ID)[0]
    assert k == src.size and k > 1000
    got = _np_rays(d_out)[:k]
    o, d, t = _orig(rays)[src], _dirs(rays)[src], hits["t"][src]
    p = (o + (d * t[:, None]).astype(f32)).astype(f32)
    g = R.geometric_normals(pos, idx, hits["primitive_id"][src].astype(np.int64), oracle.mat4_inverse_or_identity(m), o, p)
    want_dir, want_dist = R.towards_point(p, light)
    # the geometric normal goes through a normalisation (rsqrt-free, but sqrt + division): allow the offset to differ by the
    # truncation of 256 * n at most by one unit in the last place
    want_o = R.offset_ray_hit(p, g)
    assert np.all(np.abs(_orig(got).view(np.int32).astype(np.int64) - want_o.view(np.int32).astype(np.int64)) <= 1)
    assert float(np.mean(_orig(got).view(np.int32) == want_o.view(np.int32))) > 0.999
    assert np.array_equal(_dirs(got), want_dir) and np.array_equal(got["tmax"], want_dist)
    assert np.all(got["tmin"] == np.finfo(f32).eps)
    # shadow test: first accepted hit ends the search (reference-order kernel), bit-identical to the oracle on these rays
    d_sh = torch.zeros_like(d_out)
    flags = api.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH
    sp.p.trace_closest_device(d_out.data_ptr(), k, d_sh.data_ptr(), ray_flags=flags, stream=st)
    torch.cuda.synchronize()
    want = sp.o.trace(got, ray_flags=flags, n_threads=4, want_counters=False)
    got_h = d_sh.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)[:k]
    assert got_h.tobytes() == want.tobytes()
    occluded = int((got_h["instance_id"] != api.INVALID_ID).sum())
    assert 0 < occluded < k  # part of the torus is in its own shadow, part is lit


def test_ao_frame_accumulates_like_the_reference_ray_gen_shader():
    """feature/ao.rs:150-330 as a device-resident pipeline, three samples: camera primary rays (CULL_BACK) -> compacted AO test rays
    (mode 1, sample index = sample count, range 0.01..100, ACCEPT_FIRST_HIT_AND_END_SEARCH) -> payload (miss 1 / occluded 0) ->
    running mean in the AO buffer.  Each wave's hits are checked against the oracle on the rays the device generated, the buffer
    against the same f32 recurrence in numpy."""
    import torch
    sp, _ = helpers.torus_scene(96)
    W, H = 160, 120
    n = W * H
    st = torch.cuda.current_stream().cuda_stream
    d_rays, d_hits, d_ao, d_aoh = _dev_rays(n), torch.zeros((n, 32), dtype=torch.uint8, device="cuda"), _dev_rays(n), torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    d_src = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_buf = torch.zeros(n, dtype=torch.float32, device="cuda")
    want_buf = np.zeros(n, f32)
    first_hit = api.RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH
    for sample in range(3):
        sp.p.gen_pinhole_rays_device(d_rays.data_ptr(), W, H, tmin=0.0, tmax=1e30, stream=st)  # (camera model is tested separately)
        sp.p.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=helpers.CULL_BACK, grid_width=W, stream=st)
        sp.p.gen_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_ao.data_ptr(), d_src.data_ptr(), d_n.data_ptr(), mode=1,
                                    sample_index=sample, max_sample=256, tmin=0.01, tmax=100.0, stream=st)
        k = int(d_n.item())
        sp.p.trace_closest_device(d_ao.data_ptr(), k, d_aoh.data_ptr(), ray_flags=first_hit, stream=st)
        sp.p.ao_accumulate_device(d_aoh.data_ptr(), d_src.data_ptr(), d_n.data_ptr(), n, sample, d_buf.data_ptr(), stream=st)
        torch.cuda.synchronize()
        prim = d_hits.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
        ao_rays = _np_rays(d_ao)[:k]
        sec = d_aoh.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)[:k]
        assert sec.tobytes() == sp.o.trace(ao_rays, ray_flags=first_hit, n_threads=4, want_counters=False).tobytes()
        src = d_src.cpu().numpy().view(np.uint32)[:k]
        assert np.array_equal(src, np.nonzero(prim["instance_id"] != api.INVALID_ID)[0])
        payload = np.ones(n, f32)
        payload[src] = np.where(sec["instance_id"] != api.INVALID_ID, f32(0.0), f32(1.0))
        want_buf = ((want_buf * f32(sample)).astype(f32) + payload).astype(f32) / f32(sample + 1)
        assert np.array_equal(d_buf.cpu().numpy(), want_buf.astype(f32)), sample
    assert 0.0 < float(want_buf[prim["instance_id"] != api.INVALID_ID].mean()) < 1.0   # the torus occludes part of its own hemisphere
    assert np.all(want_buf[prim["instance_id"] == api.INVALID_ID] == 1.0)
    # frozen once max_sample samples are in
    sp.p.ao_accumulate_device(d_aoh.data_ptr(), d_src.data_ptr(), d_n.data_ptr(), n, 256, d_buf.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(d_buf.cpu().numpy(), want_buf.astype(f32))


def test_device_resident_wavefront_matches_the_oracle_on_the_rays_it_generated():
    """primary (device gen) -> closest hit -> compacted cosine bounce (device gen) -> closest hit, no host round trip in between;
    both waves bit-identical to the oracle traversal of the same rays"""
    import torch
    sp, _ = helpers.torus_scene(128)
    W, H = 320, 200
    n = W * H
    d_rays, d_hits = _primary_and_hits(sp, W, H)
    st = torch.cuda.current_stream().cuda_stream
    d_b = _dev_rays(n)
    d_bh = torch.zeros_like(d_b)
    d_src = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    sp.p.gen_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_b.data_ptr(), d_src.data_ptr(), d_n.data_ptr(), mode=0, stream=st)
    k = int(d_n.item())  # the only host read: the wave size (the reference reads it back the same way, task_group.rs:259-277)
    sp.p.trace_closest_device(d_b.data_ptr(), k, d_bh.data_ptr(), ray_flags=0, stream=st)
    torch.cuda.synchronize()
    prim_rays, bounce_rays = _np_rays(d_rays), _np_rays(d_b)[:k]
    want1 = sp.o.trace(prim_rays, ray_flags=helpers.CULL_BACK, n_threads=4, want_counters=False)
    want2 = sp.o.trace(bounce_rays, ray_flags=0, n_threads=4, want_counters=False)
    got1 = d_hits.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
    got2 = d_bh.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)[:k]
    assert got1.tobytes() == want1.tobytes()
    assert got2.tobytes() == want2.tobytes()
    assert int((got2["instance_id"] != api.INVALID_ID).sum()) > 100  # the torus re-hits itself


@pytest.mark.parametrize("first_hit", [False, True])
def test_wave_sized_on_the_device_needs_no_host_read(first_hit):
    """rdn_rt_trace_closest_device_n: the bounce wave is launched over the upper bound with its size left on the device by the
    bounce step — no read-back between the waves (both kernels: the ordered one and, for ACCEPT_FIRST_HIT rays, the
    reference-order one).  Records beyond the device-side count stay untouched."""
    import torch
    sp, _ = helpers.torus_scene(128)
    W, H = 320, 200
    n = W * H
    d_rays, d_hits = _primary_and_hits(sp, W, H)
    st = torch.cuda.current_stream().cuda_stream
    d_b = _dev_rays(n)
    d_bh = torch.full((n, 32), 0xAB, dtype=torch.uint8, device="cuda")
    d_src = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    flags = 0x04 if first_hit else 0
    sp.p.gen_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_b.data_ptr(), d_src.data_ptr(), d_n.data_ptr(), mode=0, stream=st)
    sp.p.trace_closest_device_n(d_b.data_ptr(), d_n.data_ptr(), n, d_bh.data_ptr(), ray_flags=flags, stream=st)
    assert sp.p.poll_errors(stream=st) == 0
    k = int(d_n.item())
    assert 0 < k < n
    want = sp.o.trace(_np_rays(d_b)[:k], ray_flags=flags, n_threads=4, want_counters=False)
    got = d_bh.cpu().numpy()
    assert got[:k].view(api.HIT_DTYPE).reshape(-1).tobytes() == want.tobytes()
    assert np.all(got[k:] == 0xAB)
