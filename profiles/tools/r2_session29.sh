#!/bin/bash
# Round-2 GPU session 29 (one GPU): suite with the final defaults, captures of the one-launch traversal, bench line
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/t29.log 2>&1; tail -6 gpurun_out/t29.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
PLF_GRAPH=0 timeout 300 ncu --kernel-name regex:'k_clv|k_cherry' --metrics $M --clock-control none --csv \
  --log-file gpurun_out/r2_traffic_narrow.csv python profiles/tools/traffic_run.py narrow > gpurun_out/tr_narrow.log 2>&1
tail -2 gpurun_out/tr_narrow.log
PLF_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:'k_clv_dna_flow' -c 3 \
  -o /tmp/flow python profiles/tools/traffic_run.py narrow > gpurun_out/ncu_flow.log 2>&1
ncu -i /tmp/flow.ncu-rep --page details --csv > gpurun_out/r2_full_flow_details.csv 2>/dev/null
python profiles/tools/ncu_details.py gpurun_out/r2_full_flow_details.csv > gpurun_out/r2_full_flow_summary.txt; tail -4 gpurun_out/r2_full_flow_summary.txt
python bench.py > gpurun_out/b29.log 2>gpurun_out/b29.err; echo "bench exit code $?"; tail -c 200 gpurun_out/b29.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke29.log 2>&1; echo "smoke exit code $?"; tail -1 gpurun_out/smoke29.log
