#!/bin/bash
OUT=gpurun_out/${1:-b2b}; shift; mkdir -p $OUT
for v in "$@"; do for c in ${CFGS:-c2 c4 c1}; do
  RDN_ORDERED_VARIANT=$v python tools/kbench.py $c 40 2>&1 | tail -1 | sed 's/skip_tie=False //; s/build_s=[0-9.]* //; s/bit_identical_sample/ok/; s/with events between [0-9.]* //' >> $OUT/b2b.log
done; done
cat $OUT/b2b.log
