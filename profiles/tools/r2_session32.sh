#!/bin/bash
# Round-2 GPU session 32 (one GPU): full suite with the final path-kernel defaults, 1000-taxon thresholds, bench line, smoke
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/t32.log 2>&1; tail -5 gpurun_out/t32.log
timeout 300 python profiles/tools/narrow_kinds.py > gpurun_out/narrow_kinds5.json 2>gpurun_out/narrow_kinds5.err; echo "rc $?"
python bench.py > gpurun_out/b32.log 2>gpurun_out/b32.err; echo "bench exit code $?"; tail -c 200 gpurun_out/b32.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke32.log 2>&1; echo "smoke exit code $?"; tail -1 gpurun_out/smoke32.log
