#!/bin/bash
# Round-2 GPU session 14 (one GPU): the whole GPU suite including the guard-band suite
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t14.log 2>&1; tail -8 gpurun_out/t14.log
