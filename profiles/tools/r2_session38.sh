#!/bin/bash
# Round-2 GPU session 38 (one GPU): dependencies of plain lists under site repeats; the suite
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "repeats or plain_lists or flow") > gpurun_out/t38.log 2>&1; tail -8 gpurun_out/t38.log
