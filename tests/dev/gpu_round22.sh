#!/bin/bash
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
timeout 500 python bench.py --steps 5 --warmup 3 --no-frame > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; tail -2 gpurun_out/bench_last.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_last.json').read().strip().splitlines()[-1])
e=d['e2e']; print(d['value'], e['value'], e['ms_per_step'], e['mode'], e['one_step_at_a_time'], e['stream'])
PY
