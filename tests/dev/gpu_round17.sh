#!/bin/bash
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
LZ4B200_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-frame > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
grep "# e2e\|Error" gpurun_out/bench_stream.err; tail -3 gpurun_out/bench_stream.err | cut -c1-300
