#!/bin/bash
# `ncu --set full` captures of the tensor-core kernels from one eager step of the bench workload (run under gpurun).
set -u
mkdir -p gpurun_out
B=${B:-16}
cap() {  # name regex skip count
  timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" -s "$3" -c "$4" \
      -f -o "gpurun_out/r01_$1" python bench.py --res R2 --batch $B --steps 1 --warmup 3 --no-cpu-baseline --mode eager --ncu-step \
      > "gpurun_out/ncu_$1.log" 2>&1
  ncu -i "gpurun_out/r01_$1.ncu-rep" --page raw --csv 2>/dev/null > "gpurun_out/r01_$1_raw.csv"
  python tools/ncu_pick.py all < "gpurun_out/r01_$1_raw.csv" > "gpurun_out/r01_$1.txt" 2>&1
  tail -40 "gpurun_out/r01_$1.txt"
}
cap fwd "^tapconv_fwd_kernel$" 90 8
cap wgrad "^tapconv_wgrad_kernel$" 30 8
