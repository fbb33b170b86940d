#!/bin/bash
# Round-2 GPU session 7 (one GPU): whole GPU suite, default bench line, DRAM-byte captures of the CLV traversals,
# ncu --set full of the dominant kernels exported to text (the .ncu-rep files stay on the box: too large).
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t7.log 2>&1; tail -6 gpurun_out/t7.log
python bench.py > gpurun_out/b7.log 2>gpurun_out/b7.err; tail -c 300 gpurun_out/b7.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for c in dna aa repeats repeats_ids; do
  PLF_GRAPH=0 ncu --kernel-name regex:'k_clv|k_cherry|k_rid|k_rep_pairs' --metrics $M --clock-control none --csv \
    --log-file gpurun_out/r2_traffic_$c.csv python profiles/tools/traffic_run.py $c > gpurun_out/tr_$c.log 2>&1
  tail -1 gpurun_out/tr_$c.log
done
NCU="ncu --set full --import-source on --clock-control none"
full() { # name, kernel regex, count, traffic_run args...
  name=$1; re=$2; n=$3; shift 3
  PLF_GRAPH=0 $NCU --kernel-name regex:"$re" -c $n -f -o /tmp/$name python profiles/tools/traffic_run.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/r2_full_${name}_details.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r2_full_${name}_raw.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep gpurun_out/r2_full_${name}_*.csv
}
full dna_stream 'k_clv_dna_stream' 12 dna --sites 400000 --reps 1
full aa_stream 'k_clv_aa_mma_stream' 8 aa --reps 1
full repeats 'k_clv_dna_ii_pairs|k_rid_' 30 repeats_ids --reps 1
du -sh gpurun_out
