#!/bin/bash
# K2 CTA size (build-time variants), then the other BASELINE configs with the product library
mkdir -p gpurun_out
for w in 1 2; do bash tests/dev/build_variant.sh dw$w -DDEC_WARPS_PER_CTA=$w > /dev/null 2>&1; done
python -c "from lz4_flex_b200 import _native; print(_native.build())" > gpurun_out/build.log 2>&1
for f in compression_66k_JSON.txt dickens.txt; do
  for so in liblz4b200.so liblz4b200_dw2.so liblz4b200_dw1.so; do
    LZ4B200_SO_OVERRIDE=$PWD/lz4_flex_b200/$so timeout 200 python tests/dev/thread_sweep.py 16384 $f plain 2>&1 | tail -1
  done
done | tee gpurun_out/k2_cta_size.txt
timeout 600 python tests/dev/config_probe.py 2>&1 | tee gpurun_out/config_probe.txt | tail -8
