#!/bin/bash
# producer-side back-off (matchers waiting for a free ring half): 0 (spin) / 100 / 200 (default) / 500 ns; then the
# lane-group and shared-memory-tag kernels again with the back-off in place
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
for f in compression_66k_JSON.txt dickens.txt; do
  for so in liblz4b200.so liblz4b200_ps0.so liblz4b200_ps100.so liblz4b200_ps500.so; do
    LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/$so timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
  done
  timeout 300 python tests/dev/thread_sweep.py 16384 $f nib 2>&1 | grep -v oracle
  timeout 400 python tests/dev/thread_sweep.py 16384 $f g16 2>&1 | grep -v oracle
done | tee gpurun_out/sweep_backoff.txt
timeout 600 python -m pytest tests/test_gpu_block.py -x -q -m gpu 2>&1 | tail -2
