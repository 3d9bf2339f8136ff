#!/bin/bash
# Final evidence pass of round 2: whole GPU suite, smoke, both bench arms, launch list, sanitizers.
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu_final.txt 2>&1; tail -4 gpurun_out/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-260 gpurun_out/r2_bench_ref.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'c/d ms', d['compress_ms'], d['decompress_ms'], 'frac', d['roofline']['frac'], d['roofline']['traffic'], d['roofline_decompress']['frac'], d['roofline_decompress']['traffic'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('compress_call_ms'), d['e2e'].get('decompress_call_ms'), d['e2e'].get('link'), 'cpu', d['cpu_baseline'])
f = d['sharded_frame']; print('frame', f['value'], f['ms_per_step'], f['collective'], f['parity']['byte_identical_to_oracle'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-frame > gpurun_out/launches.log 2>&1
timeout 2400 bash tests/dev/sanitize2.sh > gpurun_out/sanitize.out 2>&1; tail -40 gpurun_out/sanitize_summary.txt
