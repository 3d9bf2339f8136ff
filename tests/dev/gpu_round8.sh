#!/bin/bash
# one-warp decode CTAs in the product launcher; e2e worker-pool caller candidates
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
for f in compression_66k_JSON.txt dickens.txt; do
  timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
done | tee gpurun_out/k2_onewarp.txt
timeout 600 python -m pytest tests/test_gpu_block.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_block.txt
LZ4B200_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-frame > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
grep "# e2e" gpurun_out/bench_e2e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print(d['value'], d['compress_ms'], d['decompress_ms'], d['e2e']['ms_per_step'], d['e2e']['api'], d['roofline_decompress'])
PY
