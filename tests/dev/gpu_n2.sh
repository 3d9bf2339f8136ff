#!/bin/bash
# 2-GPU validation of the sharded frame path (peer-memory gather over NVLink) and of the whole bench line at N=2
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload frame --steps 5 --warmup 2 --frame-blocks 64 > gpurun_out/frame_n2.json 2> gpurun_out/frame_n2.err
tail -3 gpurun_out/frame_n2.err; cut -c1-900 gpurun_out/frame_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
f = d['sharded_frame']; print('frame', f['value'], f['ms_per_step'], f['collective'], f['parity'], f['e2e']['value'])
PY
