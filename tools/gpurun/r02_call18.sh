#!/bin/bash
# Round 2, GPU call 18: tap-channel PatchGAN head (tests, A/B timing, launch list), k-tile rule, full suite, bench.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x ) > gpurun_out/c18_kernels.log 2>&1
tail -4 gpurun_out/c18_kernels.log | cut -c1-300
timeout 300 python tools/head_probe.py > gpurun_out/c18_head_probe.log 2>&1; cat gpurun_out/c18_head_probe.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c18_head_launches.csv python tools/head_probe.py > /dev/null 2>&1
python - <<'P'
import csv
rows = list(csv.reader(l for l in open("gpurun_out/c18_head_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
seq = [(r[ki][:70], float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0)) for r in rows[1:]]
import collections
agg = collections.OrderedDict()
for n, us in seq:
    if "at::" in n: continue
    agg.setdefault(n, []).append(us)
for n, v in agg.items():
    print("%4d x  %s   %s" % (len(v), n, " ".join("%.1f" % u for u in sorted(set(round(x, 0) for x in v))[:6])))
P
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/c18_pytest.log 2>&1
tail -4 gpurun_out/c18_pytest.log | cut -c1-300
( time timeout 1500 python bench.py --no-library-baseline ) > gpurun_out/c18_bench.log 2> gpurun_out/c18_bench.err
grep '^{' gpurun_out/c18_bench.log | head -c 600; echo; tail -3 gpurun_out/c18_bench.err
cp gpurun_out/kernel_profile_c2_R2_b16.tsv gpurun_out/c18_kernel_profile_c2_R2_b16.tsv 2>/dev/null
