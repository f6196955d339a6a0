#!/bin/bash
# compile-time variants of the compaction kernels (only compact.cu is rebuilt), then an ncu capture of the shipped configuration
#   bash tools/gpu_compact_variants.sh <tag> ["<flags>;<flags>;..."]
OUT=gpurun_out/${1:-compactv}; mkdir -p $OUT
SPECS=${2:-";-DRDN_COMPACT_MINB=6;-DRDN_COMPACT_MINB=5;-DRDN_COMPACT_CB=128 -DRDN_COMPACT_MINB=16;-DRDN_COMPACT_CB=128 -DRDN_COMPACT_MINB=12;-DRDN_COMPACT_CB=512 -DRDN_COMPACT_MINB=4;-DRDN_COMPACT_CB=512 -DRDN_COMPACT_MINB=3"}
IFS=';' read -ra FL <<< "$SPECS"
for f in "${FL[@]}"; do
  touch rendiation_b200/csrc/compact.cu
  RDN_EXTRA_NVCC_FLAGS="$f" python -m rendiation_b200.build > /dev/null 2>&1
  echo "== flags [$f]" >> $OUT/variants.log
  python tools/compact_bench.py 20 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n'], d['keep_fraction'], round(d['ms_median'], 4), 'ms', round(d['gbs']), 'GB/s', round(d['frac_of_hbm_peak'], 3), d['matches_numpy'])" >> $OUT/variants.log
done
touch rendiation_b200/csrc/compact.cu; python -m rendiation_b200.build > /dev/null 2>&1
cat $OUT/variants.log
python tools/compact_bench.py 20 > $OUT/compact.jsonl 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:k_compact -s 12 -c 3 -o $OUT/prof_compact python tools/compact_bench.py 1 > $OUT/ncu_compact.log 2>&1
ls -la $OUT
