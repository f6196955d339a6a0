#!/bin/bash
set -u
OUT=gpurun_out/${1:-c5}; mkdir -p $OUT
python tools/c5_run.py --check >> $OUT/c5.jsonl 2>> $OUT/c5.err
for n in 2 4 8; do
  for t in 512 256; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n tools/c5_run.py --check --tile $t >> $OUT/c5.jsonl 2>> $OUT/c5.err
  done
done
cut -c1-330 $OUT/c5.jsonl; grep -i "error\|Traceback" $OUT/c5.err | head
