#!/bin/bash
# Round-2 GPU session 30 (one GPU): wide shape of the path kernel: parity variants, A/B against the ring kernels, API latency
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "virtual_cherries_parity or flow") > gpurun_out/t30.log 2>&1; tail -5 gpurun_out/t30.log
timeout 600 python profiles/tools/wide_ab.py > gpurun_out/wide_ab.json 2>gpurun_out/wide_ab.err; echo "rc $?"; cat gpurun_out/wide_ab.err | tail -5
timeout 300 python profiles/tools/api_latency.py > gpurun_out/api_latency.json 2>gpurun_out/api_latency.err; echo "rc $?"
