#!/bin/bash
# Round 2, GPU call 12: swapped-operand forward mode (Cout <= 128) and merged-tap MMA in the 64-channel weight gradient.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x ) > gpurun_out/c12_kernels.log 2>&1
tail -5 gpurun_out/c12_kernels.log | cut -c1-300
SH="16 640 384 256 128 3  16 640 384 128 128 3  16 320 192 512 128 3  16 640 384 128 64 3  16 640 384 64 64 3  16 640 384 128 64 1  16 40 24 128 2048 3"
( echo "--- default (swapped operands for Cout <= 128, merged taps)"
  timeout 600 python tools/conv_probe.py $SH
  timeout 600 python tools/conv_probe.py --seg
  echo "--- dbg6=8 dbg5=2 (previous kernels)"
  timeout 600 python tools/conv_probe.py --dbg6=8 --dbg5=2 $SH
  timeout 600 python tools/conv_probe.py --dbg6=8 --seg ) > gpurun_out/c12_probe.log 2>&1
cat gpurun_out/c12_probe.log | cut -c1-200
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c12_pytest.log 2>&1
tail -6 gpurun_out/c12_pytest.log | cut -c1-300
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c12_smoke.log 2>&1
grep "smoke ok" gpurun_out/c12_smoke.log | cut -c1-200
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c12_bench.log 2> gpurun_out/c12_bench.err
grep '^{' gpurun_out/c12_bench.log | head -c 500; echo; tail -3 gpurun_out/c12_bench.err
