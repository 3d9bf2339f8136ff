#!/bin/bash
# GPU pass 3: variants tests, half-warp + persisting-L2 sweep, K1-S2 on the frame workload
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_frame.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu3.txt 2>&1
tail -8 gpurun_out/pytest_gpu3.txt
timeout 500 python tests/dev/thread_sweep.py 16384 compression_66k_JSON.txt g16 > gpurun_out/sweep_g16_json.txt 2>&1; tail -16 gpurun_out/sweep_g16_json.txt
for solo in 0 2; do
  LZ4B200_ENC_SOLO=$solo timeout 300 python bench.py --workload frame --steps 5 --warmup 2 --frame-blocks 256 2>gpurun_out/frame_solo$solo.err | tail -1 > gpurun_out/frame_solo$solo.json
  python -c "import json; d=json.load(open('gpurun_out/frame_solo$solo.json')); print('solo=$solo', d['value'], d['ms_per_step'], d['collective'], d['parity']['byte_identical_to_oracle'])"
done
