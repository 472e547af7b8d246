#!/bin/bash
# Round 2, GPU call 26: evidence on the final binaries -- per-launch ncu metrics of one whole eager step, `--set full` captures
# of the kernel classes changed in round 2e and of the dominant class, memcheck, default bench (with library baseline), reference arm.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --mode eager --ncu-step"
( time timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02f_step_metrics.csv $BENCH ) > gpurun_out/c26_stepmetrics.log 2>&1
tail -3 gpurun_out/c26_stepmetrics.log | cut -c1-200
python tools/step_metrics_summary.py gpurun_out/r02f_step_metrics.csv > gpurun_out/r02f_step_metrics_summary.txt 2>&1
head -30 gpurun_out/r02f_step_metrics_summary.txt | cut -c1-230
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --profile-from-start off -k "regex:$2" -s "$3" -c "$4" -f -o /tmp/$1 $BENCH > gpurun_out/c26_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02f_ncu_$1.csv 2>/dev/null
  python tools/ncu_pick.py all < gpurun_out/r02f_ncu_$1.csv > gpurun_out/r02f_ncu_$1.txt 2>&1
  grep "time_duration\|dram__bytes\|tensor_cycles_active.avg.pct_of_peak_sustained_active\|dram_throughput" gpurun_out/r02f_ncu_$1.txt | cut -c1-200
}
cap fused_spade "tapconv_fwd_kernel<256, 0, 2" 0 2
cap swapped "tapconv_fwd_kernel<128, 0, 3" 20 3
cap wgrad256 "tapconv_wgrad_kernel<256>" 0 4
cap spade_bwd "spade_bwd_(reduce|apply)" 0 6
cap inorm_bwd "instnorm_bwd" 0 6
cap head_img "head_dots|head_gather|head_scatter|img_dgrad|img_wgrad|thin_out1_tile" 0 8
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_kernels.py -k "conv_img or head or instance_norm or tcgen05_vs_torch or simt or multitap" ) > gpurun_out/c26_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/c26_memcheck.log | cut -c1-200
( time timeout 1500 python bench.py ) > gpurun_out/c26_bench_default.log 2> gpurun_out/c26_bench_default.err
grep '^{' gpurun_out/c26_bench_default.log | head -c 300; echo; tail -3 gpurun_out/c26_bench_default.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/c26_bench_reference.log 2> gpurun_out/c26_bench_reference.err
grep '^{' gpurun_out/c26_bench_reference.log | head -c 600; echo; tail -3 gpurun_out/c26_bench_reference.err
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c26_smi.txt 2>&1
