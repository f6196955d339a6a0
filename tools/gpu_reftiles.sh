OUT=gpurun_out/rt1; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for a in 1 0; do for c in c2 c4 c1; do
  RDN_REF_TILES=$a KBENCH_FLAGS=0x14 timeout 300 python tools/kbench.py $c 20 2>&1 | tail -1 | sed "s/^/ref_tiles=$a /; s/skip_tie=False //; s/build_s=[0-9.]* //; s/bit_identical_sample/ok/; s/all 20 results identical to the serialised one/same/; s/pdl=1 side_stream=0//" >> $OUT/ref.log
done; done
cat $OUT/ref.log
