#!/bin/bash
# timeline-only debug build (three timestamps per warp): when does the ray list run dry, when does the last warp leave
OUT=gpurun_out/${1:-tl}; mkdir -p $OUT
RDN_EXTRA_NVCC_FLAGS="-DRDN_DEBUG_TIMELINE" python -m rendiation_b200.build --force > /dev/null 2>&1
for c in c2 c3 c4 c1; do
  echo "== timeline cfg=$c" >> $OUT/timeline.log
  KBENCH_CHECK=0 python tools/kbench.py $c 3 2>&1 | grep "dbg steps\|cfg=" | tail -2 | sed 's/.*timeline:/timeline:/' >> $OUT/timeline.log
done
python -m rendiation_b200.build --force > /dev/null 2>&1
cat $OUT/timeline.log
