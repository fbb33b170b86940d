#!/bin/bash
# Round-2 GPU session 34 (one GPU): virtual cherries inside the path kernel: parity, then A/B
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_guard.py -x -q -m gpu) > gpurun_out/t34.log 2>&1; tail -5 gpurun_out/t34.log
timeout 600 python profiles/tools/narrow_ab.py > gpurun_out/narrow_flow5.json 2>gpurun_out/narrow_flow5.err; echo "rc $?"
