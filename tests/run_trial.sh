#!/bin/bash
# development aid (run under gpurun): ncu capture of K1 split with source-level sampling
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_compress -s 2 -c 1 -f -o gpurun_out/k1_split python bench.py --steps 1 --warmup 3 --quick 2>&1 | tail -3
ls -la gpurun_out
