#!/bin/bash
# Round-2 GPU session: the whole GPU test suite, then ncu DRAM-byte captures of the CLV traversals of the
# bench configurations (profiles/tools/traffic_run.py), then the config-2-shaped bench line.  Run from the
# repo root on the GPU box; everything lands in gpurun_out/.
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t2.log 2>&1; tail -12 gpurun_out/t2.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for c in ${CAPTURES:-dna aa repeats repeats_ids}; do
  PLF_GRAPH=0 ncu --kernel-name regex:'k_clv|k_cherry|k_rid|k_rep' --metrics $M --clock-control none --csv \
    --log-file gpurun_out/r2_traffic_$c.csv python profiles/tools/traffic_run.py $c > gpurun_out/tr_$c.log 2>&1
  tail -2 gpurun_out/tr_$c.log
done
python bench.py --sites 1000000 --no-configs > gpurun_out/b2_1M.log 2>&1; tail -c 1500 gpurun_out/b2_1M.log
