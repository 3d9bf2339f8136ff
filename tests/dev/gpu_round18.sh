#!/bin/bash
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
LZ4B200_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-frame > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
grep "# e2e stream\|Error" gpurun_out/bench_stream.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_stream.json').read().strip().splitlines()[-1])
print(d['e2e']['link'], d['e2e']['ms_per_step'], d['e2e']['stream'])
PY
