#!/bin/bash
# Round-2 GPU session 3: GPU test suite, the default bench line, config-2-shaped lines (A/B of the cherry tile shape)
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t3.log 2>&1; tail -8 gpurun_out/t3.log
python bench.py > gpurun_out/b3.log 2>&1; tail -c 300 gpurun_out/b3.log
python bench.py --sites 1000000 --no-configs --no-cpu-baseline > gpurun_out/b3_1M.log 2>&1
PLF_CHERRY_ITEMS=4 python bench.py --sites 1000000 --no-configs --no-cpu-baseline > gpurun_out/b3_1M_items4.log 2>&1
PLF_VIRTUAL_CHERRIES=0 python bench.py --sites 1000000 --no-configs --no-cpu-baseline > gpurun_out/b3_1M_nocherry.log 2>&1
for f in b3_1M b3_1M_items4 b3_1M_nocherry; do python - <<PY
import json
for line in open("gpurun_out/$f.log"):
    if line.startswith("{"):
        d = json.loads(line); print("$f", round(d["ms_per_step"], 4), d["step_breakdown_ms"]["clv_updates"], d["roofline"]["frac_moved"], d["roofline"]["frac"])
PY
done
