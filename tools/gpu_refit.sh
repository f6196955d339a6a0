#!/bin/bash
# TLAS-only update timings with the stage breakdown (RDN_BUILD_TIMING), pool on / off, subtree task sizes
#   bash tools/gpu_refit.sh <tag>
OUT=gpurun_out/${1:-refit}; mkdir -p $OUT
RDN_BUILD_TIMING=1 python tools/refit_bench.py 20 > $OUT/refit.json 2> $OUT/refit_stages.log
tail -4 $OUT/refit_stages.log; cat $OUT/refit.json
for mt in 512 2048 4096; do RDN_BUILD_MIN_TASK=$mt python tools/refit_bench.py 20 2>/dev/null | sed "s/^/min_task $mt: /" | tee -a $OUT/refit_variants.log; done
RDN_BUILD_POOL=0 python tools/refit_bench.py 20 2>/dev/null | sed 's/^/pool off: /' | tee -a $OUT/refit_variants.log
