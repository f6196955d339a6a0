#!/bin/bash
OUT=gpurun_out/${1:-compact}; mkdir -p $OUT
python tools/compact_bench.py 20 > $OUT/compact.jsonl 2> $OUT/compact.err
timeout 600 ncu --set full --clock-control none -k regex:k_compact_u32 -s 4 -c 2 -o $OUT/prof_compact python tools/compact_bench.py 1 > $OUT/ncu_compact.log 2>&1
cat $OUT/compact.jsonl; tail -2 $OUT/compact.err
