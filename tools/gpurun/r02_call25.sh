#!/bin/bash
# Round 2, GPU call 25: shared SPADE statistics + tile-kernel occupancy: suite, smoke, bench.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c25_pytest.log 2>&1
tail -4 gpurun_out/c25_pytest.log | cut -c1-300
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c25_smoke.log 2>&1
grep "smoke ok" gpurun_out/c25_smoke.log | cut -c1-300
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c25_bench.log 2> gpurun_out/c25_bench.err
grep '^{' gpurun_out/c25_bench.log | head -c 400; echo; tail -3 gpurun_out/c25_bench.err
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c25_kernel_profile_c2_R2_b16.tsv 2>/dev/null
