#!/bin/bash
# timeline-only debug build (timestamps per warp and per tile pass): when does the ray list run dry, when does the last warp leave,
# how long do the passes over a tile take.   Usage: [CFGS="c2 c3"] bash tools/gpu_timeline.sh <tag> [variants...]
OUT=gpurun_out/${1:-tl}; shift; mkdir -p $OUT
RDN_EXTRA_NVCC_FLAGS="-DRDN_DEBUG_TIMELINE" python -m rendiation_b200.build --force > /dev/null 2>&1
for v in ${@:-0}; do for c in ${CFGS:-c2 c3 c4 c1}; do
  echo "== timeline cfg=$c variant=$v" >> $OUT/timeline.log
  RDN_ORDERED_VARIANT=$v KBENCH_CHECK=0 python tools/kbench.py $c 3 2>&1 | grep "dbg steps\|dbg tiles" | tail -2 | sed 's/.*timeline:/timeline:/' >> $OUT/timeline.log
done; done
python -m rendiation_b200.build --force > /dev/null 2>&1
cat $OUT/timeline.log
