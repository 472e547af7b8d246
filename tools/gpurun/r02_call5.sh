#!/bin/bash
# Round 2, GPU call 5: suite, conv probe (multi-lane TMA issue in the wgrad kernels), bench.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/c5_pytest.log 2>&1
tail -15 gpurun_out/c5_pytest.log
SHAPES="16 640 384 64 64 3 16 640 384 128 64 3 16 640 384 128 128 3 16 640 384 256 128 3 16 320 192 512 128 3 16 640 384 128 256 3 16 320 192 128 512 3 16 40 24 1024 1024 3"
( timeout 300 python tools/conv_probe.py $SHAPES ) > gpurun_out/c5_probe.log 2>&1
cat gpurun_out/c5_probe.log
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c5_bench.log 2> gpurun_out/c5_bench.err
head -c 1500 gpurun_out/c5_bench.log; tail -5 gpurun_out/c5_bench.err
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c5_kernel_profile_c2_R2_b16.tsv 2>/dev/null
