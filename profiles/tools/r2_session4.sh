#!/bin/bash
# Round-2 GPU session 4: A/B of the config-2-shaped bench line against the round-1 build of the library
# (profiles/tools/_ab/libpll_b200_r1.so, built from commit 6fbde77) in one session, interleaved; the new
# 20-state / repeats tests; ncu --set full captures of the dominant kernels.
mkdir -p gpurun_out
B="python bench.py --sites 1000000 --no-configs --no-cpu-baseline"
for round in 1 2; do
  PLL_B200_LIB=$PWD/profiles/tools/_ab/libpll_b200_r1.so $B > gpurun_out/ab_r1_$round.log 2>&1
  PLF_VIRTUAL_CHERRIES=0 $B > gpurun_out/ab_new_nocherry_$round.log 2>&1
  $B > gpurun_out/ab_new_$round.log 2>&1
done
for f in gpurun_out/ab_*.log; do python - <<PY
import json
for line in open("$f"):
    if line.startswith("{"):
        d = json.loads(line); print("$f", round(d["ms_per_step"], 4), round(d["step_breakdown_ms"]["clv_updates"] if "step_breakdown_ms" in d else d["roofline"]["ms_per_step_in_kernel"], 4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
(time python -m pytest tests/test_gpu_round2.py -x -q -m gpu) > gpurun_out/t4.log 2>&1; tail -5 gpurun_out/t4.log
NCU="ncu --set full --import-source on --clock-control none"
PLF_GRAPH=0 $NCU --kernel-name regex:k_clv_dna_stream -c 14 -f -o gpurun_out/r2_full_dna_stream python profiles/tools/traffic_run.py dna --reps 1 > gpurun_out/ncu_dna.log 2>&1
PLF_GRAPH=0 $NCU --kernel-name regex:k_clv_aa_mma_stream -c 8 -f -o gpurun_out/r2_full_aa_stream python profiles/tools/traffic_run.py aa --reps 1 > gpurun_out/ncu_aa.log 2>&1
PLF_GRAPH=0 $NCU --kernel-name regex:'k_clv_dna_ii_pairs|k_rid_' -c 40 -f -o gpurun_out/r2_full_repeats python profiles/tools/traffic_run.py repeats_ids --reps 1 > gpurun_out/ncu_rep.log 2>&1
ls -la gpurun_out/*.ncu-rep
