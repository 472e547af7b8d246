#!/bin/bash
# Round 2, GPU call 37: the rebuilt library after the cosmetic edit: kernel tests of conv_thin.cu + smoke.
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_img or head or simt" ) > gpurun_out/c37_kernels.log 2>&1
tail -2 gpurun_out/c37_kernels.log | cut -c1-200
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c37_smoke.log 2>&1
grep "smoke ok" gpurun_out/c37_smoke.log | cut -c1-120
