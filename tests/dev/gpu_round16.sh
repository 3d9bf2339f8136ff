#!/bin/bash
# e2e stream of batches under pipeline variants
mkdir -p gpurun_out
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
run() { echo "## $*"; env "$@" LZ4B200_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-frame 2>&1 >/dev/null | grep "# e2e stream\|Error\|error" ; }
run LZ4B200_PIPE_GTAB=0
run LZ4B200_PIPE_GTAB=1
run LZ4B200_PIPE_GTAB=1 LZ4B200_DEC_CHUNK_MB=64
run LZ4B200_PIPE_GTAB=1 LZ4B200_ENC_CHUNK_MB=64
run LZ4B200_PIPE_GTAB=1 LZ4B200_ENC_CHUNK_MB=256
