#!/bin/bash
# Round-2 GPU session 23 (one GPU): one launch per level on narrow alignments: A/B over widths + the tests that cover it
mkdir -p gpurun_out
python profiles/tools/narrow_ab.py > gpurun_out/narrow_level.json 2>gpurun_out/narrow.err; python - <<PY
import json
d = json.load(open("gpurun_out/narrow_level.json"))
for k, v in d.items(): print(k, {n: v[n] for n in v if n.endswith("_us")}, len({v[n] for n in v if n.endswith("_logl")}) == 1)
PY
(time python -m pytest tests/test_gpu_round2.py tests/test_gpu_guard.py -x -q -m gpu) > gpurun_out/t23.log 2>&1; tail -6 gpurun_out/t23.log
