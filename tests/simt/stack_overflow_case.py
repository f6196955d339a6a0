"""TEST INFRASTRUCTURE: with the traversal stack of the emulated kernel cut to a few entries (RDN_SIMT_DEFINES=RDN_STACK_MAX=3),
a trace must FAIL with "traversal stack overflow" — through the host-buffer call and through the device-resident call with
statistics — instead of returning records with dropped subtrees.  (The real stack holds 120 entries for trees of depth 50 + 50;
no scene the builder produces can overflow it, so the report is unreachable on a GPU.)"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)
os.environ["RDN_SIMT_DEFINES"] = "RDN_STACK_MAX=3"
import numpy as np

import build_emu
from rendiation_b200 import api, scenes as S

api.LIB_PATH = build_emu.build()
import helpers

sp, _ = helpers.torus_scene(96, product=True)
rays = S.pinhole_rays(128, 96, 0.01, 100.0)
try:
    sp.p.trace_closest_batch(rays, ray_flags=0x10, grid_width=128)
except api.RdnError as e:
    assert "traversal stack overflow" in str(e), e
else:
    raise SystemExit("the host-buffer trace returned although the traversal stack overflowed")


def aligned(n, dtype):
    raw = np.zeros(n * 32 + 64, np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + n * 32].view(dtype)


d_rays, d_hits = aligned(rays.shape[0], api.RAY_DTYPE), aligned(rays.shape[0], api.HIT_DTYPE)
d_rays[:] = rays
try:
    sp.p.trace_closest_device(d_rays.ctypes.data, rays.shape[0], d_hits.ctypes.data, ray_flags=0x10, grid_width=128, want_stats=True)
except api.RdnError as e:
    assert "traversal stack overflow" in str(e), e
else:
    raise SystemExit("the device-resident trace with statistics returned although the traversal stack overflowed")
# the reference-order kernel is stackless: the same rays through it are the oracle's
got, _ = sp.p.trace_counted(rays, ray_flags=0x10)
want, _ = sp.o.trace(rays, ray_flags=0x10, n_threads=4)
assert got.tobytes() == want.tobytes()
print("stack overflow is reported, the stackless kernel is unaffected")
