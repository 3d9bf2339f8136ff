#!/bin/bash
# first GPU pass of round 2: the whole GPU suite (no -x: new kernel families are tested side by side), then the
# launch-shape sweeps; everything lands in gpurun_out/
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
timeout 400 python tests/dev/thread_sweep.py 16384 > gpurun_out/sweep_json.txt 2>&1; tail -12 gpurun_out/sweep_json.txt
timeout 300 python tests/dev/thread_sweep.py 16384 dickens.txt > gpurun_out/sweep_dickens.txt 2>&1; tail -10 gpurun_out/sweep_dickens.txt
timeout 200 python bench.py --workload frame --steps 3 --warmup 1 --frame-blocks 64 2>gpurun_out/frame64.err | tail -1 > gpurun_out/frame64.json; cat gpurun_out/frame64.json | cut -c1-400
