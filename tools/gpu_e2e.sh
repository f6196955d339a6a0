#!/bin/bash
# host-buffer path, page-locked and pageable caller buffers: staging threads, streaming stores, chunk sizes   bash tools/gpu_e2e.sh <tag>
OUT=gpurun_out/${1:-e2e}; mkdir -p $OUT
python tools/e2e_bench.py 20 2>&1 | tail -2 | tee -a $OUT/e2e.log
RDN_STAGE_STREAMING=0 python tools/e2e_bench.py 20 2>&1 | tail -1 | sed 's/^/no streaming stores: /' | tee -a $OUT/e2e.log
for t in 4 8 12; do RDN_STAGE_THREADS=$t python tools/e2e_bench.py 20 2>&1 | tail -1 | tee -a $OUT/e2e.log; done
for c in 65536 131072 524288; do RDN_HOST_CHUNK_RAYS=$c python tools/e2e_bench.py 20 2>&1 | tail -1 | tee -a $OUT/e2e.log; done
