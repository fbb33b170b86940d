#!/bin/bash
# Round-2 GPU session 15 (one GPU): ring depth / CTA width A/B of the 20-state streaming kernel on config 3
mkdir -p gpurun_out
for v in "" "PLF_AA_STAGES=6" "PLF_AA_WARPS=8" "PLF_AA_STAGES=6 PLF_AA_WARPS=8" "PLF_VIRTUAL_CHERRIES=0" "PLF_VIRTUAL_CHERRIES=0 PLF_AA_STAGES=6"; do
  env $v python profiles/tools/config3_quick.py 2>>gpurun_out/c3.err | tee -a gpurun_out/c3_ab.jsonl
done
