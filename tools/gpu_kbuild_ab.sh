#!/bin/bash
# A/B of compile-time knobs of the ordered kernel (rebuilds the library on the box per setting)   bash tools/gpu_kbuild_ab.sh <tag> "<flags>;<flags>;..."
OUT=gpurun_out/${1:-kb}; mkdir -p $OUT
IFS=';' read -ra SPECS <<< "$2"
for f in "${SPECS[@]}"; do
  touch rendiation_b200/csrc/traverse.cu; RDN_EXTRA_NVCC_FLAGS="$f" python -m rendiation_b200.build > /dev/null 2>&1
  for c in ${CFGS:-c2 c3 c4 c1}; do
    echo "== [$f] $c $(timeout 300 python tools/kbench.py $c ${ITERS:-30} 2>&1 | tail -1 | sed 's/.*mean_ms/mean_ms/; s/, all [0-9]* results identical to the serialised one//; s/pdl=1 side_stream=0//; s/bit_identical_sample/ok/; s/build_s=[0-9.]* //')" >> $OUT/kbuild.log
  done
done
touch rendiation_b200/csrc/traverse.cu; python -m rendiation_b200.build > /dev/null 2>&1
cat $OUT/kbuild.log
