#!/bin/bash
# Round-2 GPU session 21 (one GPU): suite with the memoised identifier updates, config-4 timings, final repeats captures
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t21.log 2>&1; tail -6 gpurun_out/t21.log
python profiles/tools/config4_quick.py > gpurun_out/c4_memo.json 2>gpurun_out/c4.err; cat gpurun_out/c4_memo.json
PLL_CUDA_REPEATS_MEMO=0 python profiles/tools/config4_quick.py > gpurun_out/c4_nomemo.json 2>>gpurun_out/c4.err; cat gpurun_out/c4_nomemo.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for c in repeats repeats_ids; do
  PLF_GRAPH=0 ncu --kernel-name regex:'k_clv|k_cherry|k_rid|k_rep_pairs' --metrics $M --clock-control none --csv \
    --log-file gpurun_out/r2_traffic_$c.csv python profiles/tools/traffic_run.py $c > gpurun_out/tr_$c.log 2>&1
  tail -1 gpurun_out/tr_$c.log
done
