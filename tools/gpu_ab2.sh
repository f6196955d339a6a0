#!/bin/bash
# A/B of ordered-kernel variants without the test-suite: bash tools/gpu_ab2.sh <tag> <variants...>   (CFGS, ITERS from the environment)
OUT=gpurun_out/${1:-ab}; shift; mkdir -p $OUT
for c in ${CFGS:-c2 c3 c4 c1}; do for v in "$@"; do
  echo "== $c v$v $(RDN_ORDERED_VARIANT=$v timeout 300 python tools/kbench.py $c ${ITERS:-30} 2>&1 | tail -1 | sed 's/skip_tie=False //; s/build_s=[0-9.]* //; s/bit_identical_sample/ok/; s/irregular=0\/0 //; s/, all [0-9]* results identical to the serialised one//; s/pdl=1 side_stream=0//')" >> $OUT/ab.log
done; done
cat $OUT/ab.log
