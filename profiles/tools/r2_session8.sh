#!/bin/bash
# Round-2 GPU session 8 (one GPU): suite, A/B of the cherry-consumer variants on the config-2 shape, default bench
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t8.log 2>&1; tail -6 gpurun_out/t8.log
B="python bench.py --sites 1000000 --no-configs --no-cpu-baseline"
for round in 1 2; do
  $B > gpurun_out/ab8_default_$round.log 2>&1
  PLF_CHERRY_STAGES=4 $B > gpurun_out/ab8_stages4_$round.log 2>&1
  PLF_CHERRY_BULK=0 $B > gpurun_out/ab8_ring_$round.log 2>&1
done
for f in gpurun_out/ab8_*.log; do python - <<PY
import json
for line in open("$f"):
    if line.startswith("{"):
        d = json.loads(line); print("$f", round(d["ms_per_step"], 4), round(d["step_breakdown_ms"]["clv_updates"], 4), round(d["roofline"]["frac_moved"], 4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
python bench.py > gpurun_out/b8.log 2>gpurun_out/b8.err; tail -c 200 gpurun_out/b8.log
