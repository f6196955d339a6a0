#!/bin/bash
# GPU tests, kbench of the default kernels, bench (both arms), ncu --set full of the shipped kernel on configs 1-4 + launch list
set -u
TAG=${1:-r2i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
CFGS="c2 c3 c4 c1" bash tools/gpu_ab3.sh $TAG "0;100;110" > /dev/null 2>&1
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
python bench.py --config c5 --steps 3 --warmup 1 > $OUT/bench_c5.json 2>> $OUT/bench.err
if [ "${NCU:-1}" = "1" ]; then
for c in c1 c2 c3 c4; do
  KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_ordered -s 3 -c 1 \
      -o $OUT/prof_ordered_$c python tools/kbench.py $c 3 > $OUT/ncu_full_$c.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --c5-steps 0 --other-configs 0 > $OUT/bench_under_ncu.log 2>&1
fi
tail -3 $OUT/pytest_gpu.log; cat $OUT/ab.log; cat $OUT/bench.json; tail -3 $OUT/bench.err; cat $OUT/bench_c5.json
