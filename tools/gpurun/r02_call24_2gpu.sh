#!/bin/bash
# Round 2, GPU call 24 (2 GPUs): two-rank parity test and the N = 2 bench line on the round-2e kernels; c4 / c5 lines at N = 1.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_multirank.py -q -x ) > gpurun_out/c24_multirank.log 2>&1
tail -4 gpurun_out/c24_multirank.log | cut -c1-300
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/c24_bench2_c2.log 2> gpurun_out/c24_bench2_c2.err
grep '^{' gpurun_out/c24_bench2_c2.log | head -c 500; echo; tail -3 gpurun_out/c24_bench2_c2.err
for WL in c4 c5; do
  ( time timeout 600 python bench.py --workload $WL --no-library-baseline ) > gpurun_out/c24_bench_$WL.log 2> gpurun_out/c24_bench_$WL.err
  grep '^{' gpurun_out/c24_bench_$WL.log | head -c 500; echo; tail -3 gpurun_out/c24_bench_$WL.err
done
