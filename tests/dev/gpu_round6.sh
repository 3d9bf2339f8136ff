#!/bin/bash
# K2 compressed-stream prefetch (build-time variants) vs the product library
mkdir -p gpurun_out
for d in 128 256 512; do bash tests/dev/build_variant.sh pf$d -DDEC_PREFETCH=$d > /dev/null 2>&1; done
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
for f in compression_66k_JSON.txt dickens.txt; do
  for so in liblz4b200.so liblz4b200_pf128.so liblz4b200_pf256.so liblz4b200_pf512.so; do
    LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/$so timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
  done
done | tee gpurun_out/k2_prefetch.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/bench_ref.err; python -c "
import json; r=json.loads(open('gpurun_out/r2_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['compress_mibs'], r['cpu_baseline']['decompress_mibs'])"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'c/d ms', d['compress_ms'], d['decompress_ms'], 'frac', d['roofline']['frac'], d['roofline']['kernel'], d['roofline']['traffic'], d['roofline_decompress']['kernel'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['compress_mibs'], d['cpu_baseline']['decompress_mibs'], d['cpu_baseline']['single_thread'], d['cpu_baseline']['liblz4_anchor'])
f = d['sharded_frame']; print('frame', f['value'], f['ms_per_step'], f['collective']['exchange_and_pack_ms'], f['parity']['byte_identical_to_oracle'], f['cpu_baseline'])
PY
