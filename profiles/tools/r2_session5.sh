#!/bin/bash
# Round-2 GPU session 5 (N GPUs of one box): the strong-scaling bench line at N ranks, the reference arm, and the
# two-rank sharded parity test over NCCL.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N > gpurun_out/b5_n$N.log 2>gpurun_out/b5_n$N.err; echo "torchrun exit code $?"; tail -c 2500 gpurun_out/b5_n$N.log; tail -3 gpurun_out/b5_n$N.err
if [ "$N" = "2" ]; then
  $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/b5_ref_n$N.log 2>&1; tail -c 600 gpurun_out/b5_ref_n$N.log
  python -m pytest tests/test_gpu_round2.py -k sharded -x -q 2>&1 | tail -3
fi
