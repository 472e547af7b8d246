#!/bin/bash
# Round 2, GPU call 9: suite + smoke + bench after the precise-math build, eval-mode backward, chsum guard.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/c9_pytest.log 2>&1
tail -12 gpurun_out/c9_pytest.log | cut -c1-300
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c9_smoke.log 2>&1
grep "smoke ok" gpurun_out/c9_smoke.log | cut -c1-200
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c9_bench.log 2> gpurun_out/c9_bench.err
grep '^{' gpurun_out/c9_bench.log | head -c 700; echo; tail -3 gpurun_out/c9_bench.err
