#!/bin/bash
# Round-2 GPU session 31 (one GPU): 1 / 2 / 4 blocks per thread in the path kernel: parity, then A/B
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_guard.py -x -q -m gpu -k "virtual_cherries_parity or flow or guard") > gpurun_out/t31.log 2>&1; tail -5 gpurun_out/t31.log
timeout 600 python profiles/tools/narrow_ab.py > gpurun_out/narrow_flow4.json 2>gpurun_out/narrow_flow4.err; echo "rc $?"
