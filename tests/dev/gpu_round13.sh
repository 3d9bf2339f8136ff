#!/bin/bash
# validation of the new defaults (gtab without the L2 hint + opaque table offset, one-warp decode CTAs, ramp-down chunk plan):
# whole GPU suite, smoke, both bench arms, launch list, ncu --set full of K1 and K2; predicated-table-access variant
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu_final.txt 2>&1; tail -6 gpurun_out/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/r2_bench_ref.json
LZ4B200_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err; grep "# e2e" gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'c/d ms', d['compress_ms'], d['decompress_ms'], 'frac', d['roofline']['frac'], d['roofline']['kernel'], d['roofline_decompress']['kernel'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('compress_call_ms'), d['e2e'].get('decompress_call_ms'), d['e2e'].get('link'), 'cpu', d['cpu_baseline']['value'])
f = d['sharded_frame']; print('frame', f['value'], f['ms_per_step'], f['collective'], f['parity']['byte_identical_to_oracle'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-frame > gpurun_out/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_blocks -s 2 -c 1 -o gpurun_out/r2_k1 python bench.py --quick --steps 2 --warmup 1 --no-frame > gpurun_out/ncu_k1.log 2>&1; tail -1 gpurun_out/ncu_k1.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_blocks -s 2 -c 1 -o gpurun_out/r2_k2 python bench.py --quick --steps 2 --warmup 1 --no-frame > gpurun_out/ncu_k2.log 2>&1; tail -1 gpurun_out/ncu_k2.log
for f in compression_66k_JSON.txt dickens.txt; do
  for so in liblz4b200.so liblz4b200_pr1.so; do
    LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/$so timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
  done
done | tee gpurun_out/sweep_pr1.txt
