#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_modules.py -q -k trajectory ) > gpurun_out/c10_traj.log 2>&1
tail -8 gpurun_out/c10_traj.log | cut -c1-400
python - <<'PY'
import json
d = json.load(open('gpurun_out/trajectory_parity.json'))
for i, (a, b) in enumerate(zip(d['ours'], d['oracle'])):
    print(i, ' '.join('%s %.4f/%.4f' % (k.split('/')[0][:6], a[k], b[k]) for k in a))
PY
