#!/bin/bash
# Round-2 GPU session 28 (one GPU): k_clv_dna_flow with register-carried paths: suite, then A/B
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_guard.py -x -q -m gpu) > gpurun_out/t28a.log 2>&1; tail -8 gpurun_out/t28a.log
(time timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_round2.py --deselect tests/test_gpu_guard.py) > gpurun_out/t28b.log 2>&1; tail -5 gpurun_out/t28b.log
timeout 600 python profiles/tools/narrow_ab.py > gpurun_out/narrow_flow3.json 2>gpurun_out/narrow_flow3.err; echo "rc $?"
timeout 300 python profiles/tools/narrow_kinds.py > gpurun_out/narrow_kinds4.json 2>gpurun_out/narrow_kinds4.err; echo "rc $?"
python bench.py --no-cpu-baseline > gpurun_out/b28.log 2>gpurun_out/b28.err; echo "bench exit code $?"; tail -c 300 gpurun_out/b28.log
