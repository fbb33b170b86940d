#!/bin/bash
# Round-2 GPU session 42 (one GPU): the final build - whole suite in one process, bench line, smoke
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/t42.log 2>&1; tail -5 gpurun_out/t42.log
python bench.py > gpurun_out/b42.log 2>gpurun_out/b42.err; echo "bench exit code $?"; tail -c 200 gpurun_out/b42.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke42.log 2>&1; echo "smoke exit code $?"; tail -1 gpurun_out/smoke42.log
