#!/bin/bash
OUT=gpurun_out/${1:-compactv}; mkdir -p $OUT
for f in "" "-DRDN_COMPACT_MINB=8" "-DRDN_COMPACT_CB=128" "-DRDN_COMPACT_CB=128 -DRDN_COMPACT_MINB=12" "-DRDN_COMPACT_STATIC_TILES" "-DRDN_COMPACT_CB=128 -DRDN_COMPACT_MINB=12 -DRDN_COMPACT_STATIC_TILES" "-DRDN_COMPACT_CB=512"; do
  RDN_EXTRA_NVCC_FLAGS="$f" python -m rendiation_b200.build --force > /dev/null 2>&1
  echo "== flags [$f]" >> $OUT/variants.log
  python tools/compact_bench.py 20 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n'], d['keep_fraction'], round(d['ms_median'], 4), 'ms', round(d['gbs']), 'GB/s', round(d['frac_of_hbm_peak'], 3), d['matches_numpy'])" >> $OUT/variants.log
done
python -m rendiation_b200.build --force > /dev/null 2>&1
cat $OUT/variants.log
