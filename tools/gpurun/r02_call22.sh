#!/bin/bash
# Round 2, GPU call 22: rewritten InstanceNorm backward / forward-apply kernels: tests, suite, bench, ncu of the IN kernels.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "instance_norm" ) > gpurun_out/c22_in.log 2>&1
tail -5 gpurun_out/c22_in.log | cut -c1-300
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c22_pytest.log 2>&1
tail -4 gpurun_out/c22_pytest.log | cut -c1-300
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c22_bench.log 2> gpurun_out/c22_bench.err
grep '^{' gpurun_out/c22_bench.log | head -c 400; echo; tail -3 gpurun_out/c22_bench.err
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c22_kernel_profile_c2_R2_b16.tsv 2>/dev/null
timeout 700 ncu --set full --clock-control none --profile-from-start off -k "regex:instnorm" -c 12 -f -o /tmp/inorm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --mode eager --ncu-step > gpurun_out/c22_ncu_inorm.log 2>&1
ncu -i /tmp/inorm.ncu-rep --page raw --csv > gpurun_out/r02e_ncu_inorm.csv 2>/dev/null
python tools/ncu_pick.py all < gpurun_out/r02e_ncu_inorm.csv > gpurun_out/r02e_ncu_inorm.txt 2>&1
grep "Kernel Name\|time_duration\|dram__bytes\|dram_throughput\|grid_size" gpurun_out/r02e_ncu_inorm.txt | cut -c1-400
