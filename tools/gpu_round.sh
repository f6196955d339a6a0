#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernel.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.csv 2>&1
nproc > $OUT/nproc.txt
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 > $OUT/bench_under_ncu.log 2>&1
KBENCH_CHECK=0 ncu --set full --clock-control none --import-source on -k regex:k_trace_ordered -s 3 -c 2 \
    -o $OUT/prof_ordered python tools/kbench.py c2 3 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/bench.json; tail -2 $OUT/bench.err
