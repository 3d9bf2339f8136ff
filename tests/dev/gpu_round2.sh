#!/bin/bash
# GPU pass 2: fixed tests, tagged-table sweep, one ncu capture of the tagged kernel, the whole bench line
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.txt 2>&1
tail -12 gpurun_out/pytest_gpu.txt
timeout 400 python tests/dev/thread_sweep.py 16384 compression_66k_JSON.txt gtag > gpurun_out/sweep_gtag_json.txt 2>&1; tail -9 gpurun_out/sweep_gtag_json.txt
timeout 300 python tests/dev/thread_sweep.py 16384 dickens.txt gtag > gpurun_out/sweep_gtag_dickens.txt 2>&1; tail -9 gpurun_out/sweep_gtag_dickens.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err; cut -c1-1500 gpurun_out/bench_r2a.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_blocks_gtag -s 2 -c 1 -o gpurun_out/r2_k1_gtag python bench.py --quick --steps 2 --warmup 1 --no-frame > gpurun_out/ncu_gtag.log 2>&1; tail -2 gpurun_out/ncu_gtag.log
