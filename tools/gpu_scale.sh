#!/bin/bash
# 8-GPU session: bench.py at N = 1, 2, 4, 8 and BASELINE configs[4] (tools/c5_run.py) at N = 8 with three sharding tile sizes.
set -u
OUT=gpurun_out/${1:-scale}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.csv 2>&1
python bench.py --steps 50 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 \
    > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
done
for t in 512 256 128; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/c5_run.py --check --tile $t \
    >> $OUT/c5_n8.jsonl 2>> $OUT/c5_n8.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 tools/c5_run.py --tile 128 >> $OUT/c5_n4.jsonl 2>> $OUT/c5_n4.err
for n in 1 2 4 8; do python - <<PY
import json
d=json.loads(open("$OUT/bench_n$n.json").read().strip().splitlines()[-1])
print("N=$n value=%.0f e2e=%.0f ms_per_step=%.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
done
cat $OUT/c5_n8.jsonl $OUT/c5_n4.jsonl | cut -c1-400
