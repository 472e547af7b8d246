#!/bin/bash
# Round 2, GPU call 34: GAN loss sums fused into the logit head's gather kernel: tests, suite, bench.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gan_loss_sums or head" ) > gpurun_out/c34_gan.log 2>&1
tail -12 gpurun_out/c34_gan.log | cut -c1-300
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c34_pytest.log 2>&1
tail -4 gpurun_out/c34_pytest.log | cut -c1-300
( time timeout 1500 python bench.py --no-library-baseline --no-cpu-baseline ) > gpurun_out/c34_bench.log 2> gpurun_out/c34_bench.err
grep '^{' gpurun_out/c34_bench.log | head -c 300; echo; tail -3 gpurun_out/c34_bench.err
