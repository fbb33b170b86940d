#!/bin/bash
# Round-2 GPU session 11 (one GPU): narrow-alignment A/B of virtual cherries, 256-thread cherry consumers A/B,
# the GPU suite, pair-list kernel under ncu --set full
mkdir -p gpurun_out
python profiles/tools/narrow_ab.py > gpurun_out/narrow_ab.json 2>gpurun_out/narrow_ab.err; cat gpurun_out/narrow_ab.json
B="python bench.py --sites 1000000 --no-configs --no-cpu-baseline"
for round in 1 2; do
  $B > gpurun_out/ab11_default_$round.log 2>&1
  PLF_CHERRY_ITEMS=1 $B > gpurun_out/ab11_threads256_$round.log 2>&1
done
for f in gpurun_out/ab11_*.log; do python - <<PY
import json
for line in open("$f"):
    if line.startswith("{"):
        d = json.loads(line); print("$f", round(d["ms_per_step"], 4), round(d["step_breakdown_ms"]["clv_updates"], 4), round(d["roofline"]["frac_moved"], 4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t11.log 2>&1; tail -6 gpurun_out/t11.log
NCU="ncu --set full --import-source on --clock-control none"
PLF_GRAPH=0 $NCU --kernel-name regex:'k_clv_dna_ii_pairs' --launch-skip 3 -c 5 -f -o /tmp/pairs python profiles/tools/traffic_run.py repeats --reps 1 > gpurun_out/ncu_pairs.log 2>&1
ncu -i /tmp/pairs.ncu-rep --page details --csv > gpurun_out/r2_full_pairs_details.csv 2>/dev/null
ncu -i /tmp/pairs.ncu-rep --page raw --csv > gpurun_out/r2_full_pairs_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_pairs_*
