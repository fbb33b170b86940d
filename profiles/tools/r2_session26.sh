#!/bin/bash
# Round-2 GPU session 26 (one GPU): k_clv_dna_flow (whole narrow traversal in one launch): suite, then A/B
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/t26.log 2>&1; tail -8 gpurun_out/t26.log
timeout 600 python profiles/tools/narrow_ab.py > gpurun_out/narrow_flow.json 2>gpurun_out/narrow_flow.err; echo "rc $?"
timeout 300 python profiles/tools/narrow_kinds.py > gpurun_out/narrow_kinds2.json 2>gpurun_out/narrow_kinds2.err; echo "rc $?"
