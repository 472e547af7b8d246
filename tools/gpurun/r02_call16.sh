#!/bin/bash
# Round 2, GPU call 16: tiled PatchGAN-head kernel (tests + A/B timing), k-tile experiments for the D weight gradients.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "simt or head" ) > gpurun_out/c16_kernels.log 2>&1
tail -4 gpurun_out/c16_kernels.log | cut -c1-300
timeout 300 python tools/head_probe.py > gpurun_out/c16_head_probe.log 2>&1; cat gpurun_out/c16_head_probe.log | cut -c1-200
for kt in "" "16,4,1" "8,2,4" "2,2,16" "4,2,8" "8,8,1"; do
  echo "--- S2E_KTILE=$kt"
  S2E_KTILE=$kt timeout 300 python tools/conv_probe.py 32 81 49 256 512 4 32 41 25 256 512 4 32 161 97 256 128 2 2>&1 | tail -3
done > gpurun_out/c16_ktile_probe.log 2>&1
cat gpurun_out/c16_ktile_probe.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:tapconv_wgrad -c 1 -o /tmp/dwgrad python tools/conv_probe.py --once 32 81 49 256 512 4 > gpurun_out/c16_ncu.log 2>&1
ncu -i /tmp/dwgrad.ncu-rep --page raw --csv > gpurun_out/r02e_ncu_dwgrad_256to512_k4.csv 2>/dev/null
python tools/ncu_pick.py all < gpurun_out/r02e_ncu_dwgrad_256to512_k4.csv 2>&1 | cut -c1-900
