#!/bin/bash
# Round 2, GPU call 14: kernel-variant tests, memcheck of the swapped-operand kernels, ncu captures, suite, smoke, bench.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x ) > gpurun_out/c14_kernels.log 2>&1
tail -4 gpurun_out/c14_kernels.log | cut -c1-300
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x \
    "tests/test_gpu_kernels.py::test_conv_tcgen05_forward_kernel_variants" "tests/test_gpu_kernels.py::test_conv_tcgen05_vs_torch" \
    "tests/test_gpu_kernels.py::test_conv_tcgen05_multitap_wgrad" "tests/test_gpu_kernels.py::test_conv_residual_in_epilogue" \
    "tests/test_gpu_kernels.py::test_relu_backward_fused_into_dgrad_epilogue" ) > gpurun_out/c14_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/c14_memcheck.log | cut -c1-200
cap() {
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:$rx -c $cnt -o /tmp/$name "$@" > gpurun_out/c14_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02d_ncu_$name.csv 2>/dev/null
}
cap swapped256to128 'tapconv_' 3 python tools/conv_probe.py --once 16 640 384 256 128 3
cap swapped128to64 'tapconv_' 3 python tools/conv_probe.py --once 16 640 384 128 64 3
cap wgradmt128 'tapconv_wgrad' 1 python tools/conv_probe.py --once 16 640 384 128 128 3
cap segfwd 'tapconv_fwd' 1 python tools/conv_probe.py --once --seg
ls -la gpurun_out/r02d_ncu_*.csv
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c14_pytest.log 2>&1
tail -4 gpurun_out/c14_pytest.log | cut -c1-300
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c14_smoke.log 2>&1
grep "smoke ok" gpurun_out/c14_smoke.log | cut -c1-200
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c14_bench.log 2> gpurun_out/c14_bench.err
grep '^{' gpurun_out/c14_bench.log | head -c 400; echo; tail -3 gpurun_out/c14_bench.err
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c14_kernel_profile_c2_R2_b16.tsv 2>/dev/null
