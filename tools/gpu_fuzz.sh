#!/bin/bash
# Fuzz parity session: debug listing of all three profiles, the GPU test-suite, quick kernel timings.
OUT=gpurun_out/${1:-fz}; mkdir -p $OUT
timeout 600 python tools/fuzz_debug.py > $OUT/fuzz_debug.log 2>&1; tail -1 $OUT/fuzz_debug.log
grep -c "ordered_mismatch=0" $OUT/fuzz_debug.log; grep -v "ordered_mismatch=0$" $OUT/fuzz_debug.log | head -40
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for c in c2 c4 c1; do timeout 300 python tools/kbench.py $c 40 2>&1 | tail -1 >> $OUT/kbench.log; done; cat $OUT/kbench.log
