#!/bin/bash
# Round-2 GPU session 24 (one GPU): the suite with the narrow-alignment defaults, default bench line
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t24.log 2>&1; tail -6 gpurun_out/t24.log
python bench.py > gpurun_out/b24.log 2>gpurun_out/b24.err; echo "bench exit code $?"; tail -c 200 gpurun_out/b24.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke24.log 2>&1; echo "smoke exit code $?"; tail -1 gpurun_out/smoke24.log
