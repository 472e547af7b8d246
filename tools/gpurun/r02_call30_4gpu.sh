#!/bin/bash
# Round 2, GPU call 30 (4 GPUs): c2 under torchrun at N = 4 on the final kernels.
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 4 --steps 5 --warmup 3 ) > gpurun_out/c30_bench4_c2.log 2> gpurun_out/c30_bench4_c2.err
grep '^{' gpurun_out/c30_bench4_c2.log | head -c 500; echo; tail -3 gpurun_out/c30_bench4_c2.err
