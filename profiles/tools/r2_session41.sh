#!/bin/bash
# Round-2 GPU session 41 (one GPU): the narrow 20-state test and the tests after it
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_guard.py tests/test_gpu_parity.py -x -q -m gpu -k "aa or AA or guard") > gpurun_out/t41.log 2>&1; tail -6 gpurun_out/t41.log
