#!/bin/bash
# Round-2 GPU session 44 (one GPU): the 20-state thresholds - whole suite, narrow kinds
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/t44.log 2>&1; tail -5 gpurun_out/t44.log
timeout 300 python profiles/tools/narrow_kinds.py > gpurun_out/narrow_kinds8.json 2>gpurun_out/narrow_kinds8.err; echo "rc $?"
