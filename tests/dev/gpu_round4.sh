#!/bin/bash
# GPU pass 4: lane-group matcher sweep (G = 16, 8) on JSON and dickens, ncu of two shapes, variant tests from the A/B library
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
timeout 500 python tests/dev/thread_sweep.py 16384 compression_66k_JSON.txt g16 > gpurun_out/sweep_g_json.txt 2>&1; tail -14 gpurun_out/sweep_g_json.txt
timeout 500 python tests/dev/thread_sweep.py 16384 dickens.txt g16 > gpurun_out/sweep_g_dickens.txt 2>&1; tail -14 gpurun_out/sweep_g_dickens.txt
for shape in 71 871; do
  LZ4B200_ENC_G16=$shape timeout 300 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_blocks_gtabg -s 2 -c 1 -o gpurun_out/r2_k1_g$shape python bench.py --quick --steps 2 --warmup 1 --no-frame > gpurun_out/ncu_g$shape.log 2>&1; tail -1 gpurun_out/ncu_g$shape.log
done
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_variants.txt 2>&1; tail -15 gpurun_out/pytest_variants.txt
# two-deep match deferral in K2 (build-time variant)
bash tests/dev/build_variant.sh defer2 -DDEC_DEFER2=1 > /dev/null 2>&1
for f in compression_66k_JSON.txt dickens.txt; do
  LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/liblz4b200.so timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
  LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/liblz4b200_defer2.so timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
done | tee gpurun_out/defer2.txt
