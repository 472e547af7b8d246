#!/bin/bash
# Round 2, GPU call 35: feature-matching L1 fused into the InstanceNorm apply kernel: tests, suite, bench.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "feature_matching or instance_norm or gan_loss_sums" ) > gpurun_out/c35_fm.log 2>&1
tail -12 gpurun_out/c35_fm.log | cut -c1-300
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c35_pytest.log 2>&1
tail -4 gpurun_out/c35_pytest.log | cut -c1-300
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c35_bench.log 2> gpurun_out/c35_bench.err
grep '^{' gpurun_out/c35_bench.log | head -c 300; echo; tail -3 gpurun_out/c35_bench.err
