#!/bin/bash
# A/B of ordered-kernel variants: GPU tests first, then kbench per (variant, config).  Usage: bash tools/gpu_ab.sh <tag> <variants...>
OUT=gpurun_out/${1:-ab}; shift; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for v in "$@"; do for c in ${CFGS:-c2 c4 c1 c3}; do
  RDN_ORDERED_VARIANT=$v timeout 300 python tools/kbench.py $c ${ITERS:-40} 2>&1 | tail -1 | sed 's/skip_tie=False //; s/build_s=[0-9.]* //; s/bit_identical_sample/ok/' >> $OUT/ab.log
done; done
cat $OUT/ab.log
