#!/bin/bash
# Round-2 GPU session 12 (one GPU): the suite with the new defaults, config-4 A/B of the pair-list kernel variants
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t12.log 2>&1; tail -6 gpurun_out/t12.log
python profiles/tools/config4_quick.py > gpurun_out/c4_default.json 2>gpurun_out/c4.err; cat gpurun_out/c4_default.json
PLF_PAIRS_CTAS=6 python profiles/tools/config4_quick.py > gpurun_out/c4_ctas6.json 2>>gpurun_out/c4.err; cat gpurun_out/c4_ctas6.json
PLL_CUDA_NO_PAIR_LISTS=1 python profiles/tools/config4_quick.py > gpurun_out/c4_nopairs.json 2>>gpurun_out/c4.err; cat gpurun_out/c4_nopairs.json
python bench.py --sites 1000000 --no-configs --no-cpu-baseline > gpurun_out/b12_1M.log 2>&1
python - <<PY
import json
for line in open("gpurun_out/b12_1M.log"):
    if line.startswith("{"):
        d = json.loads(line); print("b12_1M", round(d["ms_per_step"], 4), round(d["step_breakdown_ms"]["clv_updates"], 4), round(d["roofline"]["frac_moved"], 4), d["clocks"])
PY
