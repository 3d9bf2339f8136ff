#!/bin/bash
# 2-GPU validation of the multi-rank bench paths (run under gpurun --gpus 2)
mkdir -p gpurun_out
nvidia-smi -L
echo "== blocks workload, 2 ranks"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.json; cut -c1-700 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
echo "== frame workload, 2 ranks (NCCL gather to rank 0)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload frame --frame-blocks 32 --steps 3 --warmup 3 2>gpurun_out/bench_frame_n2.err | tail -1 > gpurun_out/bench_frame_n2.json; cut -c1-700 gpurun_out/bench_frame_n2.json; tail -3 gpurun_out/bench_frame_n2.err
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
