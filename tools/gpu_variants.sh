#!/bin/bash
# A/B of the ordered-kernel variants on the BASELINE configs (kbench: L2 flushed between iterations, parity-checked sample)
set -u
TAG=${1:-v}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "$@"; do
  for c in c2 c3 c4 c1; do
    RDN_ORDERED_VARIANT=$v python tools/kbench.py $c 30 2>&1 | tail -1 >> $OUT/variants.log
  done
done
cat $OUT/variants.log
