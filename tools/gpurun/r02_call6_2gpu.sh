#!/bin/bash
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_multirank.py -q -x ) > gpurun_out/c6_multirank.log 2>&1
tail -30 gpurun_out/c6_multirank.log | cut -c1-300
( time timeout 300 python -m pytest tests/test_gpu_data.py -q ) > gpurun_out/c6_data.log 2>&1
tail -15 gpurun_out/c6_data.log | cut -c1-300
