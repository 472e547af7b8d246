#!/bin/bash
# Round 2, GPU call 32 (2 GPUs): two-rank parity test and the N = 2 bench line with the 128 MB gradient buckets.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_multirank.py -q -x ) > gpurun_out/c32_multirank.log 2>&1
tail -3 gpurun_out/c32_multirank.log | cut -c1-300
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/c32_bench2_c2.log 2> gpurun_out/c32_bench2_c2.err
grep '^{' gpurun_out/c32_bench2_c2.log | head -c 300; echo; tail -3 gpurun_out/c32_bench2_c2.err
