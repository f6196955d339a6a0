#!/bin/bash
OUT=gpurun_out/${1:-pdl}; mkdir -p $OUT
for ss in 1; do for pdl in 1 0; do for c in c2 c1 c4; do
  KBENCH_SIDE_STREAM=$ss RDN_PDL=$pdl python tools/kbench.py $c 40 2>&1 | tail -1 | sed 's/skip_tie=False //; s/build_s=[0-9.]* //' >> $OUT/pdl.log
done; done; done
cat $OUT/pdl.log
