#!/bin/bash
# Round 2, GPU call 7 (8 GPUs): two-rank parity test, then c2 and c5 under torchrun at N = 8 (graph steps + NCCL between them).
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_multirank.py -q -x ) > gpurun_out/c7_multirank.log 2>&1
tail -5 gpurun_out/c7_multirank.log | cut -c1-300
for WL in c2 c5; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus 8 --steps 5 --warmup 3 --workload $WL ) > gpurun_out/c7_bench8_$WL.log 2> gpurun_out/c7_bench8_$WL.err
  grep '^{' gpurun_out/c7_bench8_$WL.log | head -c 1200; echo; tail -3 gpurun_out/c7_bench8_$WL.err
done
