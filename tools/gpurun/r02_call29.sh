#!/bin/bash
# Round 2, GPU call 29: pivot-shifted statistics: tests, suite, smoke, bench.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "statistics or instance_norm or spade_style" ) > gpurun_out/c29_stats.log 2>&1
tail -6 gpurun_out/c29_stats.log | cut -c1-300
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c29_pytest.log 2>&1
tail -4 gpurun_out/c29_pytest.log | cut -c1-300
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c29_smoke.log 2>&1
grep "smoke ok" gpurun_out/c29_smoke.log | cut -c1-200
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c29_bench.log 2> gpurun_out/c29_bench.err
grep '^{' gpurun_out/c29_bench.log | head -c 400; echo; tail -3 gpurun_out/c29_bench.err
