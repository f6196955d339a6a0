#!/bin/bash
# Round-2 first GPU session: GPU tests, A/B of the variants that were only modelled so far, ncu --set full of the shipped
# ordered kernel on configs 1-4, bench.   Usage (on the GPU box): bash tools/gpu_r2a.sh <tag>
set -u
TAG=${1:-r2a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.csv 2>&1
nproc > $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
kb() { # env... -- cfg
  timeout 300 python tools/kbench.py $1 ${ITERS:-30} 2>&1 | tail -1 | sed 's/skip_tie=False //; s/build_s=[0-9.]* //; s/bit_identical_sample/ok/' ; }
for c in c2 c4 c1 c3; do for v in 0 70 90 91 80; do
  echo "== $c v$v $(RDN_ORDERED_VARIANT=$v kb $c)" >> $OUT/ab.log
done; done
for v in 0 91; do echo "== c4 coarse-tlas v$v $(RDN_FINE_TLAS=0 RDN_ORDERED_VARIANT=$v kb c4)" >> $OUT/ab.log; done
echo "== c3 list-topup $(RDN_LIST_TOPUP=1 kb c3)" >> $OUT/ab.log
for c in c1 c2 c3 c4; do
  KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_ordered -s 3 -c 1 \
      -o $OUT/prof_ordered_$c python tools/kbench.py $c 3 > $OUT/ncu_full_$c.log 2>&1
done
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
tail -3 $OUT/pytest_gpu.log; cat $OUT/ab.log; cat $OUT/bench.json
