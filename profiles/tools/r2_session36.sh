#!/bin/bash
# Round-2 GPU session 36 (one GPU): final build - full suite, path-kernel A/B table, bench line, smoke
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/t36.log 2>&1; tail -5 gpurun_out/t36.log
timeout 600 python profiles/tools/narrow_ab.py > gpurun_out/narrow_flow7.json 2>gpurun_out/narrow_flow7.err; echo "rc $?"
python bench.py > gpurun_out/b36.log 2>gpurun_out/b36.err; echo "bench exit code $?"; tail -c 200 gpurun_out/b36.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke36.log 2>&1; echo "smoke exit code $?"; tail -1 gpurun_out/smoke36.log
