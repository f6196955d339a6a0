#!/bin/bash
# last session of a round: GPU tests, smoke, ncu --set full of the shipped kernels on configs 1-4, bench (both arms)   bash tools/gpu_final.sh <tag>
set -u
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1
for c in c1 c2 c3 c4; do
  KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_ordered -s 3 -c 1 \
      -o $OUT/prof_ordered_$c python tools/kbench.py $c 3 > $OUT/ncu_full_$c.log 2>&1
done
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench.err
python bench.py > $OUT/bench.json 2>> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
tail -2 $OUT/pytest_gpu.log; tail -1 $OUT/smoke.log; tail -2 $OUT/bench.err; cut -c1-400 $OUT/bench.json
