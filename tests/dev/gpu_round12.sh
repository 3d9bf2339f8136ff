#!/bin/bash
# gtagg: lane-group matchers + 2-bit shared-memory tags; 32 tuples per hand-off (sb32, default path)
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/liblz4b200_ab.so timeout 900 python -m pytest tests/variants_impl.py -q -p no:cacheprovider --tb=short -m gpu -k "tagg" 2>&1 | tail -5 | tee gpurun_out/pytest_tagg.txt
for f in compression_66k_JSON.txt dickens.txt; do
  timeout 400 python tests/dev/thread_sweep.py 16384 $f nib 2>&1 | grep -v oracle
  LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/liblz4b200_sb32.so timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
done | tee gpurun_out/sweep_tagg.txt
LZ4B200_ENC_NIB=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_blocks -s 2 -c 1 -o gpurun_out/r2_k1_tagg8 python bench.py --quick --steps 2 --warmup 1 --no-frame > gpurun_out/ncu_k1_tagg8.log 2>&1; tail -1 gpurun_out/ncu_k1_tagg8.log
