#!/bin/bash
# Round 2, GPU call 3: suite after the SPADE kernel rewrite + fixes, norm probe, bench (c2 with baselines).
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/c3_pytest.log 2>&1
tail -30 gpurun_out/c3_pytest.log
( timeout 300 python tools/norm_probe.py ) > gpurun_out/c3_norm_probe.log 2>&1
( timeout 300 python tools/norm_probe.py --fused 16 640 384 128 1 16 640 384 64 0 ) >> gpurun_out/c3_norm_probe.log 2>&1
cat gpurun_out/c3_norm_probe.log
cap() {
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:$rx -c $cnt -o /tmp/$name "$@" > gpurun_out/c3_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_$name.csv 2>/dev/null
}
cap norm 'spade_|stats_kernel' 8 python tools/norm_probe.py --once 16 640 384 128 1
cap normC64 'spade_|stats_kernel' 6 python tools/norm_probe.py --once 16 640 384 64 0
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c3_bench.log 2> gpurun_out/c3_bench.err
tail -c 5000 gpurun_out/c3_bench.log; tail -5 gpurun_out/c3_bench.err
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c3_kernel_profile_c2_R2_b16.tsv 2>/dev/null
