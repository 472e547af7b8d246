#!/bin/bash
# Round 2, GPU call 23: image-head kernels on mma.sync: tests (short timeout), memcheck, suite, bench, launch times.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_img or simt" ) > gpurun_out/c23_img.log 2>&1
tail -12 gpurun_out/c23_img.log | cut -c1-300
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_kernels.py -k "conv_img or head_conv or instance_norm" ) > gpurun_out/c23_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/c23_memcheck.log | cut -c1-200
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c23_pytest.log 2>&1
tail -4 gpurun_out/c23_pytest.log | cut -c1-300
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c23_bench.log 2> gpurun_out/c23_bench.err
grep '^{' gpurun_out/c23_bench.log | head -c 400; echo; tail -3 gpurun_out/c23_bench.err
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c23_kernel_profile_c2_R2_b16.tsv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c23_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --mode eager --ncu-step > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/c23_launches.csv 60 | grep -i "launches\|thin\|img_\|head_\|instnorm\|tanh" | cut -c1-120
