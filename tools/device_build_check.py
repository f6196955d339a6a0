"""f3: the device SAH build (csrc/build_device.cu) against the oracle / host builder, node for node, plus timing.
Usage: python tools/device_build_check.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from rendiation_b200 import api, scenes as S


def boxes_of(pos, idx):
    tri = idx.reshape(-1, 3)
    return np.concatenate([pos[tri].min(1), pos[tri].max(1)], 1).astype(np.float32)


def same(a, b):
    return all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


out = []
rng = np.random.default_rng(4)
c = rng.uniform(-1, 1, (5000, 3)).astype(np.float32); r = rng.uniform(0.01, 0.3, (5000, 3)).astype(np.float32)
cases = [("torus 48x30", boxes_of(*S.torus_mesh(48, 30))), ("sphere 40x24 (pole fans: fallback)", boxes_of(*S.uv_sphere_mesh(40, 24))),
         ("random overlapping boxes", np.concatenate([c - r, c + r], 1)), ("9 identical boxes", np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (9, 1))),
         ("torus 170x130 (44k)", boxes_of(*S.torus_mesh(170, 130))), ("torus 708x708 (1M)", boxes_of(*S.torus_mesh(708, 708)))]
for name, boxes in cases:
    for opt in ((50, 2), (10, 50)):
        t0 = time.perf_counter(); dev = api.FlattenBVH(boxes, api.SAH(4), api.TreeBuildOption(*opt), device=0); t_dev = time.perf_counter() - t0
        t0 = time.perf_counter(); host = api.FlattenBVH(boxes, api.SAH(4), api.TreeBuildOption(*opt)); t_host = time.perf_counter() - t0
        ok = same(dev.nodes, host.nodes) and np.array_equal(dev.sorted_primitive_index, host.sorted_primitive_index)
        ok_oracle = None
        if boxes.shape[0] < 100000:
            ob = oracle.FlattenBVH(boxes, oracle.STRATEGY_SAH, 4, *opt)
            hn, on = host.nodes, ob.nodes
            ok_oracle = bool(np.array_equal(ob.sorted_primitive_index, dev.sorted_primitive_index) and on.shape == dev.nodes.shape)
        out.append({"case": name, "option": opt, "primitives": int(boxes.shape[0]), "nodes": int(dev.nodes.shape[0]), "built_on_device": dev.built_on_device,
                    "equals_host_tree": bool(ok), "order_equals_oracle": ok_oracle, "device_ms": t_dev * 1e3, "host_ms": t_host * 1e3})
        print(json.dumps(out[-1]))
t0 = time.perf_counter(); api.FlattenBVH(cases[-1][1], api.SAH(4), api.TreeBuildOption(50, 2), device=0); print("second 1M device build ms", (time.perf_counter() - t0) * 1e3)
