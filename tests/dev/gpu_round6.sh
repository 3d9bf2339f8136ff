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
