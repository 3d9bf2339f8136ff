#!/bin/bash
# development aid (run under gpurun)
timeout 600 python -m pytest tests/test_gpu_frame.py -x -q 2>&1 | tail -4
