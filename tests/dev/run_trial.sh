#!/bin/bash
# development aid (run under gpurun): chunk ramp e2e, W=16 first round, then the round-end style validation
mkdir -p gpurun_out
echo "== e2e with chunk ramp"; timeout 300 python tests/dev/e2e_probe.py 2>&1 | tail -4 | head -1
echo "== parity w16"; LZ4B200_SO_OVERRIDE=$PWD/build_variants/w16.so timeout 300 python tests/dev/gpu_quick.py 2>&1 | tail -1
echo "== quick w16"; LZ4B200_SO_OVERRIDE=$PWD/build_variants/w16.so timeout 300 python bench.py --steps 5 --warmup 3 --quick 2>&1 | tail -1
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_n1.json; python -c "import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['compress_ms'], d['decompress_ms'], d['e2e'])"
