#!/bin/bash
# Round 2, GPU call 2: full suite (halo + multi-tap included), conv / norm probes, ncu captures, launch list, c4 / c5 bench lines.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/c2_pytest.log 2>&1
tail -60 gpurun_out/c2_pytest.log
SHAPES="16 640 384 64 64 3 16 640 384 128 64 3 16 640 384 128 128 3 16 640 384 64 128 3 16 320 192 128 128 3 16 640 384 256 128 3 16 320 192 512 128 3 16 320 192 256 128 3 16 640 384 128 256 3"
( timeout 300 python tools/conv_probe.py $SHAPES ) > gpurun_out/c2_probe_base.log 2>&1
( timeout 300 python tools/conv_probe.py --dbg6=1 $SHAPES ) > gpurun_out/c2_probe_halo.log 2>&1
echo BASE; cat gpurun_out/c2_probe_base.log; echo HALO; cat gpurun_out/c2_probe_halo.log
( timeout 300 python tools/norm_probe.py ) > gpurun_out/c2_norm_probe.log 2>&1
( timeout 300 python tools/norm_probe.py --fused 16 640 384 128 1 16 640 384 64 0 ) >> gpurun_out/c2_norm_probe.log 2>&1
cat gpurun_out/c2_norm_probe.log
# ---- ncu --set full captures (exported to csv here; the reports stay on the box unless small)
cap() {  # name, kernel regex, count, command...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:$rx -c $cnt -o /tmp/$name "$@" > gpurun_out/c2_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$name.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep
}
cap norm 'spade_style|stats_kernel|upsample2x_bwd' 8 python tools/norm_probe.py --once 16 640 384 128 1
cap normC64 'spade_style|stats_kernel' 6 python tools/norm_probe.py --once 16 640 384 64 0
cap fused 'tapconv' 4 python tools/norm_probe.py --once --fused 16 640 384 128 1
cap conv256 'tapconv' 4 python tools/conv_probe.py --once 16 640 384 128 256 3
cap conv64 'tapconv' 4 python tools/conv_probe.py --once 16 640 384 128 64 3
cap conv64halo 'tapconv' 4 python tools/conv_probe.py --once --dbg6=1 16 640 384 128 64 3
cap inorm 'instnorm' 6 python -m pytest tests/test_gpu_kernels.py -q -k "instance_norm_fwd_bwd"
# ---- launch list of one eager step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_R2b16_step.csv \
  python bench.py --steps 1 --warmup 3 --mode eager --ncu-step --no-cpu-baseline --no-library-baseline > gpurun_out/c2_launchlist.log 2>&1
wc -l gpurun_out/r02_launches_R2b16_step.csv
( time timeout 900 python bench.py --workload c4 ) > gpurun_out/c2_bench_c4.log 2> gpurun_out/c2_bench_c4.err
tail -c 3000 gpurun_out/c2_bench_c4.log; tail -3 gpurun_out/c2_bench_c4.err
( time timeout 900 python bench.py --workload c5 --no-library-baseline ) > gpurun_out/c2_bench_c5.log 2> gpurun_out/c2_bench_c5.err
tail -c 3000 gpurun_out/c2_bench_c5.log; tail -3 gpurun_out/c2_bench_c5.err
