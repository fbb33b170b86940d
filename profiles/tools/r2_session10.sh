#!/bin/bash
# Round-2 GPU session 10 (one GPU): suite, default bench line, config-2-shaped line, smoke
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t10.log 2>&1; tail -6 gpurun_out/t10.log
python bench.py > gpurun_out/b10.log 2>gpurun_out/b10.err; tail -c 200 gpurun_out/b10.log
python bench.py --sites 1000000 --no-configs > gpurun_out/b10_1M.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke10.log 2>&1; tail -2 gpurun_out/smoke10.log
