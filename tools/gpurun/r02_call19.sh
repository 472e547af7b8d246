#!/bin/bash
# Round 2, GPU call 19: ncu --set full of the CUDA-core "thin" kernels and a few layout kernels inside one eager step.
mkdir -p gpurun_out
cap() {  # name regex count
  timeout 700 ncu --set full --clock-control none --profile-from-start off -k "regex:$2" -c "$3" -f -o /tmp/$1 \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --mode eager --ncu-step > gpurun_out/c19_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02e_ncu_$1.csv 2>/dev/null
  python tools/ncu_pick.py all < gpurun_out/r02e_ncu_$1.csv > gpurun_out/r02e_ncu_$1.txt 2>&1
  cut -c1-260 gpurun_out/r02e_ncu_$1.txt
}
cap thin "thin_" 12
cap layout "s2d_kernel|d2s_kernel|seg_im2col|act_bwd|avgpool|make_d_input|wgrad_dot" 14
