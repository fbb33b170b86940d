#!/bin/bash
# Round-2 GPU session 25 (one GPU): Newton breakdown, narrow alignments of the other kinds, config-3 record with per-rate scalers
mkdir -p gpurun_out
python profiles/tools/newton_breakdown.py repeats > gpurun_out/newton_rep.json 2>gpurun_out/newton_rep.err; echo "rc $?"
python profiles/tools/newton_breakdown.py dna > gpurun_out/newton_dna.json 2>gpurun_out/newton_dna.err; echo "rc $?"
python profiles/tools/narrow_kinds.py > gpurun_out/narrow_kinds.json 2>gpurun_out/narrow_kinds.err; echo "rc $?"
python - > gpurun_out/c3_rate.json 2>gpurun_out/c3_rate.err <<'PY'
import json, os, importlib, torch
import bench
pkg = importlib.import_module("libpll-2_b200")
lib = pkg.load()
print(json.dumps(bench.config3_record(lib, torch, 0, os.cpu_count())))
PY
echo "rc $?"; tail -c 600 gpurun_out/c3_rate.json
