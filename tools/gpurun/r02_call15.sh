#!/bin/bash
# Round 2, GPU call 15: tcgen05.mma issue-rate probe (shape x operand layout).
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Iinclude -o /tmp/umma_rate_probe tools/umma_rate_probe.cu 2>&1 | tail -3
timeout 120 /tmp/umma_rate_probe > gpurun_out/umma_rate_probe.txt 2>&1; echo rc=$?
cat gpurun_out/umma_rate_probe.txt
