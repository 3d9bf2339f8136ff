#!/bin/bash
# development aid (run under gpurun)
mkdir -p gpurun_out
echo "== dict tests"; timeout 600 python -m pytest tests/test_gpu_dict.py -q 2>&1 | tail -4
echo "== quick default"; timeout 300 python bench.py --steps 5 --warmup 3 --quick 2>&1 | tail -1
for v in build_variants/*.so; do echo "== quick $v"; LZ4B200_SO_OVERRIDE=$PWD/$v timeout 300 python bench.py --steps 5 --warmup 3 --quick 2>&1 | tail -1; done
echo "== e2e default chunks (enc 128, dec 128)"; timeout 300 python tests/e2e_probe.py 2>&1 | tail -4
for d in 32 64 256; do echo "== e2e dec chunk $d"; LZ4B200_DEC_CHUNK_MB=$d timeout 300 python tests/e2e_probe.py 2>&1 | tail -4 | head -1; done
for e in 64 256; do echo "== e2e enc chunk $e"; LZ4B200_ENC_CHUNK_MB=$e timeout 300 python tests/e2e_probe.py 2>&1 | tail -4 | head -1; done
echo "== e2e no priority"; PRIO=0 timeout 300 python tests/e2e_probe.py 2>&1 | tail -4
