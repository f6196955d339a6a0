#!/bin/bash
# A/B lines "ENV=.. ENV=.. variant" per config:  bash tools/gpu_ab3.sh <tag> "<spec>;<spec>;..."   spec = "VAR=val VAR=val <variant>"
OUT=gpurun_out/${1:-ab}; mkdir -p $OUT
IFS=';' read -ra SPECS <<< "$2"
for c in ${CFGS:-c2 c3 c4 c1}; do for spec in "${SPECS[@]}"; do
  v=${spec##* }; envs=${spec% *}; [ "$envs" = "$spec" ] && envs=""
  echo "== $c [$spec] $(env $envs RDN_ORDERED_VARIANT=$v timeout 300 python tools/kbench.py $c ${ITERS:-30} 2>&1 | tail -1 | sed 's/.*mean_ms/mean_ms/; s/, all [0-9]* results identical to the serialised one//; s/pdl=1 side_stream=0//; s/bit_identical_sample/ok/; s/build_s=[0-9.]* //')" >> $OUT/ab.log
done; done
cat $OUT/ab.log
