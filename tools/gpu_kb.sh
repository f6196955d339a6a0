#!/bin/bash
# kbench per (variant, config) without the test-suite.  Usage: [CFGS="c2 c1"] bash tools/gpu_kb.sh <tag> <variants...>
OUT=gpurun_out/${1:-kb}; shift; mkdir -p $OUT
for v in "$@"; do for c in ${CFGS:-c2 c4 c1 c3}; do
  RDN_ORDERED_VARIANT=$v timeout 300 python tools/kbench.py $c ${ITERS:-40} 2>&1 | tail -1 | sed 's/skip_tie=False //; s/build_s=[0-9.]* //; s/bit_identical_sample/ok/; s/all 40 results identical to the serialised one/same/' >> $OUT/kb.log
done; done
cat $OUT/kb.log
