#!/bin/bash
# Round-2 GPU session 13 (one GPU): compute-sanitizer attempt on the new kernels, pair-list kernel shapes,
# final DRAM-byte captures, default bench line
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q \
  -k "virtual_cherries_parity and (12-97 or 40-1501-random-4-False) and default" > gpurun_out/sanitizer_cherries.log 2>&1
echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer_cherries.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q \
  -k "replay_a_graph or (site_repeats_parity and 24-400)" > gpurun_out/sanitizer_repeats.log 2>&1
echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer_repeats.log
python profiles/tools/config4_quick.py > gpurun_out/c4_26.json 2>gpurun_out/c4.err; cat gpurun_out/c4_26.json
PLF_PAIRS_SHAPE=44 python profiles/tools/config4_quick.py > gpurun_out/c4_44.json 2>>gpurun_out/c4.err; cat gpurun_out/c4_44.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for c in dna aa repeats repeats_ids; do
  PLF_GRAPH=0 ncu --kernel-name regex:'k_clv|k_cherry|k_rid|k_rep_pairs' --metrics $M --clock-control none --csv \
    --log-file gpurun_out/r2_traffic_$c.csv python profiles/tools/traffic_run.py $c > gpurun_out/tr_$c.log 2>&1
  tail -1 gpurun_out/tr_$c.log
done
python bench.py > gpurun_out/b13.log 2>gpurun_out/b13.err; tail -c 200 gpurun_out/b13.log
