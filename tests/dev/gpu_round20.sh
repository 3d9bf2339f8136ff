#!/bin/bash
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build()); print(_native.build_ab())" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_block.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_block.py -q -m gpu -p no:cacheprovider -k contexts_are_independent 2>&1 | tail -1; done
