#!/bin/bash
# Round 2, GPU call 31 (8 GPUs): gradient bucket size at N = 8 (no overlap in graph mode, so larger buckets only change NCCL efficiency).
mkdir -p gpurun_out
for mb in 512 128; do
  ( time S2E_BUCKET_MB=$mb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
      bench.py --gpus 8 --steps 8 --warmup 3 ) > gpurun_out/c31_bench8_mb$mb.log 2> gpurun_out/c31_bench8_mb$mb.err
  echo "S2E_BUCKET_MB=$mb"; grep '^{' gpurun_out/c31_bench8_mb$mb.log | head -c 260; echo; tail -2 gpurun_out/c31_bench8_mb$mb.err | cut -c1-200
done
