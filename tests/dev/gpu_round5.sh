#!/bin/bash
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
timeout 400 python tests/dev/thread_sweep.py 16384 compression_66k_JSON.txt g16 > gpurun_out/sweep_g2_json.txt 2>&1; tail -8 gpurun_out/sweep_g2_json.txt
timeout 400 python tests/dev/thread_sweep.py 16384 dickens.txt g16 > gpurun_out/sweep_g2_dickens.txt 2>&1; tail -8 gpurun_out/sweep_g2_dickens.txt
timeout 600 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_variants.txt 2>&1; tail -5 gpurun_out/pytest_variants.txt
