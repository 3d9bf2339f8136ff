#!/bin/bash
# round-end style validation (run under gpurun)
mkdir -p gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_default.json; python -c "import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['compress_ms'], d['decompress_ms'], d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'])"
