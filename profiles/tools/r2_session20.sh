#!/bin/bash
# Round-2 GPU session 20 (one GPU): default bench line (exit code checked), reference arm, smoke, suite,
# ncu --set full of the final 4-state streaming kernels
mkdir -p gpurun_out
python bench.py > gpurun_out/b20.log 2>gpurun_out/b20.err; echo "bench exit code $?"; tail -c 300 gpurun_out/b20.log; tail -3 gpurun_out/b20.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b20_ref.log 2>&1; echo "reference arm exit code $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke20.log 2>&1; echo "smoke exit code $?"; tail -1 gpurun_out/smoke20.log
(time python -m pytest tests -x -q -m gpu) > gpurun_out/t20.log 2>&1; tail -5 gpurun_out/t20.log
NCU="ncu --set full --import-source on --clock-control none"
PLF_GRAPH=0 $NCU --kernel-name regex:'k_clv_dna_stream' -c 12 -f -o /tmp/dna_stream python profiles/tools/traffic_run.py dna --sites 400000 --reps 1 > gpurun_out/ncu_dna.log 2>&1
ncu -i /tmp/dna_stream.ncu-rep --page details --csv > gpurun_out/r2_full_dna_stream_details.csv 2>/dev/null
ncu -i /tmp/dna_stream.ncu-rep --page raw --csv > gpurun_out/r2_full_dna_stream_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_dna_stream_*
