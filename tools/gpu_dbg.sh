#!/bin/bash
# kernel experiments: step counters + in-kernel timeline (debug build) for the variants in $DBGV, then A/B timing of the
# variants given as arguments.   Usage: DBGV="8 20" bash tools/gpu_dbg.sh <tag> <variant>...
set -u
OUT=gpurun_out/${1:-dbg}; shift; mkdir -p $OUT
if [ -n "${DBGV:-}" ]; then
  RDN_EXTRA_NVCC_FLAGS="-DRDN_DEBUG_STEPS" python -m rendiation_b200.build --force > /dev/null 2>&1
  for v in $DBGV; do for c in c2 c3 c4 c1; do
    echo "== debug counters variant=$v cfg=$c" >> $OUT/dbg.log
    RDN_ORDERED_VARIANT=$v KBENCH_CHECK=0 python tools/kbench.py $c 2 2>&1 | grep "dbg steps" | tail -1 >> $OUT/dbg.log
  done; done
  python -m rendiation_b200.build --force > /dev/null 2>&1
fi
for v in "$@"; do
  for c in ${CFGS:-c2 c3 c4 c1}; do
    RDN_ORDERED_VARIANT=$v python tools/kbench.py $c 30 2>&1 | tail -1 >> $OUT/dbg.log
  done
done
cat $OUT/dbg.log
