#!/bin/bash
# Round 2, GPU call 27: forward / data-gradient tile shape (128x1 strips vs 2-D tiles): time and DRAM bytes.
mkdir -p gpurun_out
for ft in "" "64,2,1" "32,4,1" "16,8,1"; do
  echo "--- S2E_FTILE=$ft"
  S2E_FTILE=$ft timeout 300 python tools/conv_probe.py 16 640 384 256 128 3 16 640 384 128 256 3 16 320 192 512 128 3 16 640 384 128 64 3 2>&1 | tail -4
done > gpurun_out/c27_ftile_probe.log 2>&1
cat gpurun_out/c27_ftile_probe.log | cut -c1-200
for ft in "" "32,4,1"; do
  S2E_FTILE=$ft timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tapconv_fwd -c 2 python tools/conv_probe.py --once 16 640 384 256 128 3 2>&1 | grep -v "^==" | grep "tapconv\|dram__\|duration\|hit_rate\|tensor" | cut -c1-150
done
