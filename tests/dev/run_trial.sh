#!/bin/bash
# development aid (run under gpurun): NVML clock sampler check
mkdir -p gpurun_out
timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json; python -c "import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; tail -3 gpurun_out/bench_default.err
