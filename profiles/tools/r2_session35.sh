#!/bin/bash
# Round-2 GPU session 35 (one GPU): path kernel with two blocks per thread held to 64 registers (8 CTAs per SM)
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "flow or virtual_cherries_parity or per_kind") > gpurun_out/t35.log 2>&1; tail -4 gpurun_out/t35.log
timeout 600 python profiles/tools/narrow_ab.py > gpurun_out/narrow_flow6.json 2>gpurun_out/narrow_flow6.err; echo "rc $?"
