#!/bin/bash
# sweep of the work-sharing knobs (RDN_SHARE=busy,min,after) of RDN_ORDERED_VARIANT=110   bash tools/gpu_share_sweep.sh <tag> "<knobs> ..."
OUT=gpurun_out/${1:-share}; mkdir -p $OUT
for c in ${CFGS:-c2 c3 c4}; do
  echo "== $c v91 $(RDN_ORDERED_VARIANT=91 timeout 300 python tools/kbench.py $c ${ITERS:-30} 2>&1 | tail -1 | sed 's/.*mean_ms/mean_ms/; s/, all [0-9]* results identical to the serialised one//; s/pdl=1 side_stream=0//')" >> $OUT/share.log
  for k in $2; do
  echo "== $c share=$k $(RDN_SHARE=$k RDN_ORDERED_VARIANT=110 timeout 300 python tools/kbench.py $c ${ITERS:-30} 2>&1 | tail -1 | sed 's/.*mean_ms/mean_ms/; s/, all [0-9]* results identical to the serialised one//; s/pdl=1 side_stream=0//')" >> $OUT/share.log
done; done
cat $OUT/share.log
