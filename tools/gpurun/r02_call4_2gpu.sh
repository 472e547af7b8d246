#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): multi-rank parity test, then the bench under torchrun with graph replay and eager.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c4_smi.txt
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -q -x ) > gpurun_out/c4_multirank.log 2>&1
tail -30 gpurun_out/c4_multirank.log
for MODE in graph eager; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 5 --warmup 3 --mode $MODE ) > gpurun_out/c4_bench2_$MODE.log 2> gpurun_out/c4_bench2_$MODE.err
  tail -c 2500 gpurun_out/c4_bench2_$MODE.log; tail -5 gpurun_out/c4_bench2_$MODE.err
done
