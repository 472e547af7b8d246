#!/bin/bash
# Round 2, GPU call 8: final suite, smoke, memcheck on the new kernels, final ncu captures + launch list, bench lines.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/c8_pytest.log 2>&1
tail -6 gpurun_out/c8_pytest.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c8_smoke.log 2>&1
tail -3 gpurun_out/c8_smoke.log
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_data.py tests/test_gpu_losses_tail.py \
    "tests/test_gpu_kernels.py::test_spade_style_fwd_bwd" "tests/test_gpu_kernels.py::test_spade_style_on_upsampled_input_without_materialising_it" \
    "tests/test_gpu_kernels.py::test_conv_tcgen05_halo_kernel" "tests/test_gpu_kernels.py::test_conv_tcgen05_multitap_wgrad" \
    "tests/test_gpu_kernels.py::test_spade_conv_fused_training_forward_backward" tests/test_gpu_spade35.py ) > gpurun_out/c8_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/c8_memcheck.log
cap() {
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:$rx -c $cnt -o /tmp/$name "$@" > gpurun_out/c8_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_$name.csv 2>/dev/null
}
cap norm 'spade_|stats_kernel' 8 python tools/norm_probe.py --once 16 640 384 128 1
cap normC64 'spade_|stats_kernel' 6 python tools/norm_probe.py --once 16 640 384 64 0
cap wgradmt 'tapconv_wgrad' 2 python tools/conv_probe.py --once 16 640 384 128 128 3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches_R2b16_step.csv \
  python bench.py --steps 1 --warmup 3 --mode eager --ncu-step --no-cpu-baseline --no-library-baseline > gpurun_out/c8_launchlist.log 2>&1
wc -l gpurun_out/r02c_launches_R2b16_step.csv
( timeout 300 python tools/norm_probe.py ) > gpurun_out/c8_norm_probe.log 2>&1; cat gpurun_out/c8_norm_probe.log
( time timeout 1500 python bench.py ) > gpurun_out/c8_bench.log 2> gpurun_out/c8_bench.err
grep '^{' gpurun_out/c8_bench.log | head -c 1000; echo; tail -3 gpurun_out/c8_bench.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/c8_bench_ref.log 2>&1; grep '^{' gpurun_out/c8_bench_ref.log | head -c 600; echo
( time timeout 900 python bench.py --workload c4 ) > gpurun_out/c8_bench_c4.log 2> gpurun_out/c8_bench_c4.err; grep '^{' gpurun_out/c8_bench_c4.log | head -c 500; echo
( time timeout 900 python bench.py --workload c5 ) > gpurun_out/c8_bench_c5.log 2> gpurun_out/c8_bench_c5.err; grep '^{' gpurun_out/c8_bench_c5.log | head -c 500; echo
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c8_kernel_profile_c2_R2_b16.tsv 2>/dev/null
