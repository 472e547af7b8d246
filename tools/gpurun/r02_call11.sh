#!/bin/bash
# Round 2, GPU call 11: lean forward epilogue (side-input mode at compile time, shared-space ld/st) + single tiles on small maps.
mkdir -p gpurun_out
( timeout 600 python tools/conv_probe.py --seg
  timeout 600 python tools/conv_probe.py 16 640 384 128 64 1  16 40 24 128 2048 3  16 20 12 128 2048 3  16 80 48 128 2048 3  16 640 384 64 64 3  16 640 384 128 256 3
  echo "--- dbg6=4 (double tiles everywhere)"
  timeout 600 python tools/conv_probe.py --dbg6=4 16 40 24 128 2048 3  16 20 12 128 2048 3  16 80 48 128 2048 3 ) > gpurun_out/c11_probe.log 2>&1
cat gpurun_out/c11_probe.log | cut -c1-200
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c11_pytest.log 2>&1
tail -12 gpurun_out/c11_pytest.log | cut -c1-300
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c11_smoke.log 2>&1
grep "smoke ok" gpurun_out/c11_smoke.log | cut -c1-200
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c11_bench.log 2> gpurun_out/c11_bench.err
grep '^{' gpurun_out/c11_bench.log | head -c 500; echo; tail -3 gpurun_out/c11_bench.err
