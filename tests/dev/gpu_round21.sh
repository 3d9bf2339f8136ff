#!/bin/bash
# last sanity pass of the round on the final tree: whole GPU suite, smoke, the bench line
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu_final.txt 2>&1; tail -3 gpurun_out/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'c/d ms', d['compress_ms'], d['decompress_ms'], 'frac', d['roofline']['frac'])
e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['mode'], e['one_step_at_a_time'], e['link'])
print('cpu', d['cpu_baseline']['value'])
f = d['sharded_frame']; print('frame', f['value'], f['ms_per_step'], f['parity']['byte_identical_to_oracle'])
PY
