#!/bin/bash
# tuple push through an opaque shared address (ENC_PUSH_SA) vs the default
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
for f in compression_66k_JSON.txt dickens.txt; do
  for so in liblz4b200.so liblz4b200_sa1.so; do
    LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/$so timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
  done
done | tee gpurun_out/sweep_sa1.txt
