#!/bin/bash
# Round 2, GPU call 36 (2 GPUs): two-rank parity test on the final state (fused GAN / feature-matching sums inside split graphs).
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_multirank.py -q -x ) > gpurun_out/c36_multirank.log 2>&1
tail -3 gpurun_out/c36_multirank.log | cut -c1-300
