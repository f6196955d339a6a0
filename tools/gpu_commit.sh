#!/bin/bash
# commit timings: configs[1] (1 M triangles) full commit with the phase breakdown, config 4 full commit and TLAS-only update   bash tools/gpu_commit.sh <tag>
OUT=gpurun_out/${1:-commit}; mkdir -p $OUT
RDN_BUILD_TIMING=1 python - > $OUT/commit_c2.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
from rendiation_b200 import api, scenes as S
pos, idx = S.torus_mesh(708, 708, 1.0, 0.35)
m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
for rep in range(4):
    s = api.NaiveSahBVHSystem()
    b = s.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
    s.bind_tlas([s.create_top_level_acceleration_structure(S.make_instance(m, b.id))])
    t0 = time.perf_counter(); s.commit(); print("configs[1] commit ms", round((time.perf_counter() - t0) * 1e3, 1), {k: round(v, 1) for k, v in s.build_stats().items() if "ms" in k}, flush=True)
    del s
PY
grep -v "TLAS of" $OUT/commit_c2.log | tail -12
python tools/refit_bench.py 20 2>/dev/null | tee $OUT/refit.json
