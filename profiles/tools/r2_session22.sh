#!/bin/bash
# Round-2 GPU session 22 (one GPU): fork-join of a level's runs on narrow alignments: A/B + suite
mkdir -p gpurun_out
python profiles/tools/narrow_ab.py > gpurun_out/narrow_fork.json 2>gpurun_out/narrow.err; python - <<PY
import json
d = json.load(open("gpurun_out/narrow_fork.json")); print("fork   ", {k: (v["written_us"], v["virtual_us"]) for k, v in d.items()})
PY
PLF_FORK_MAX_SITES=0 python profiles/tools/narrow_ab.py > gpurun_out/narrow_nofork.json 2>>gpurun_out/narrow.err; python - <<PY
import json
d = json.load(open("gpurun_out/narrow_nofork.json")); print("no fork", {k: (v["written_us"], v["virtual_us"]) for k, v in d.items()})
PY
PLF_FORK_MAX_SITES=1000000 python profiles/tools/narrow_ab.py > gpurun_out/narrow_forkall.json 2>>gpurun_out/narrow.err; python - <<PY
import json
d = json.load(open("gpurun_out/narrow_forkall.json")); print("fork all", {k: (v["written_us"], v["virtual_us"]) for k, v in d.items()})
PY
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t22.log 2>&1; tail -6 gpurun_out/t22.log
