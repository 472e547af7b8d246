#!/bin/bash
# Round 2, GPU call 33: final state -- suite, smoke, default bench.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c33_pytest.log 2>&1
tail -4 gpurun_out/c33_pytest.log | cut -c1-300
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c33_smoke.log 2>&1
grep "smoke ok" gpurun_out/c33_smoke.log | cut -c1-200
( time timeout 1500 python bench.py ) > gpurun_out/c33_bench.log 2> gpurun_out/c33_bench.err
grep '^{' gpurun_out/c33_bench.log | head -c 300; echo; tail -3 gpurun_out/c33_bench.err
