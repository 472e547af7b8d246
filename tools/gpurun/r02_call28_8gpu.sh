#!/bin/bash
# Round 2, GPU call 28 (8 GPUs): c2 under torchrun at N = 8 on the round-2e kernels (graph steps + NCCL between them).
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/c28_bench8_c2.log 2> gpurun_out/c28_bench8_c2.err
grep '^{' gpurun_out/c28_bench8_c2.log | head -c 700; echo; tail -3 gpurun_out/c28_bench8_c2.err
