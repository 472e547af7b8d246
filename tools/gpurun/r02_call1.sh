#!/bin/bash
# Round 2, GPU call 1: suite + full-size parity, multi-tap wgrad / halo-kernel validation, UMMA shifted-window probe, conv probes, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_smi.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/umma_shift_probe tools/umma_shift_probe.cu > gpurun_out/c1_probe.log 2>&1
timeout 120 /tmp/umma_shift_probe >> gpurun_out/c1_probe.log 2>&1
grep -v warning gpurun_out/c1_probe.log
( time timeout 1500 python -m pytest tests -m gpu -q -k "not multitap and not halo" ) > gpurun_out/c1_pytest.log 2>&1
tail -40 gpurun_out/c1_pytest.log
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k multitap ) > gpurun_out/c1_multitap.log 2>&1
tail -15 gpurun_out/c1_multitap.log
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k halo ) > gpurun_out/c1_halo.log 2>&1
tail -25 gpurun_out/c1_halo.log
NARROW="16 640 384 64 64 3 16 640 384 128 64 3 16 640 384 128 128 3 16 640 384 64 128 3 16 320 192 128 128 3 16 640 384 256 128 3"
( timeout 300 python tools/conv_probe.py $NARROW ) > gpurun_out/c1_probe_base.log 2>&1
( timeout 300 python tools/conv_probe.py --dbg5=3 $NARROW ) > gpurun_out/c1_probe_mt.log 2>&1
( timeout 300 python tools/conv_probe.py --dbg6=1 $NARROW ) > gpurun_out/c1_probe_halo.log 2>&1
echo BASE; cat gpurun_out/c1_probe_base.log; echo MT; cat gpurun_out/c1_probe_mt.log; echo HALO; cat gpurun_out/c1_probe_halo.log
( time timeout 1500 python bench.py ) > gpurun_out/c1_bench.log 2> gpurun_out/c1_bench.err
tail -c 7000 gpurun_out/c1_bench.log; tail -5 gpurun_out/c1_bench.err
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c1_kernel_profile_c2_R2_b16.tsv 2>/dev/null
