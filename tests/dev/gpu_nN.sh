#!/bin/bash
# usage: gpu_nN.sh <N>: the whole bench line at N GPUs (block record + sharded-frame record)
N=$1
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -4 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['numa'])
print('per_rank', json.dumps(d['e2e'].get('per_rank')))
f = d['sharded_frame']; print('frame', f['value'], f['ms_per_step'], f['collective'], f['parity'], f['e2e'])
PY
