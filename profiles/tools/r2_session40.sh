#!/bin/bash
# Round-2 GPU session 40 (one GPU): narrow 20-state alignments with tips as expanded CLVs: suite, then timings
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/t40.log 2>&1; tail -8 gpurun_out/t40.log
timeout 300 python profiles/tools/narrow_kinds.py > gpurun_out/narrow_kinds7.json 2>gpurun_out/narrow_kinds7.err; echo "rc $?"
