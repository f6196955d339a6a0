#!/bin/bash
# Multi-GPU session: parity tests, 1-GPU bench, N-GPU bench under torchrun, the other BASELINE configs through kbench.
# Usage: bash tools/gpu_multi.sh <tag> <n_gpus>
set -u
TAG=${1:-m}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.csv 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?" >> $OUT/bench_n1.err
for n in $(seq 2 $N); do
  if [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n \
      > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err; echo "rc=$?" >> $OUT/bench_n$n.err
  fi
done
for c in c1 c3 c4; do python tools/kbench.py $c 20 > $OUT/kbench_$c.log 2>&1; done
tail -2 $OUT/pytest_gpu.log; cat $OUT/bench_n1.json; tail -1 $OUT/bench_n1.err; cat $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err; cat $OUT/kbench_c*.log | tail -6
