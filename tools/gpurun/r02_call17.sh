#!/bin/bash
# Round 2, GPU call 17: per-kernel times of the two head routes (ncu launch list), new k-tile rule vs explicit tiles.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c17_head_launches.csv python tools/head_probe.py > gpurun_out/c17_head.log 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/c17_head_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
seq = [(r[ki][:90], float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0)) for r in rows[1:]]
# one iteration of each route = the launches between two forward kernels; print the last iteration of every (shape, route)
n = len(seq)
print(n, "launches")
for name, us in seq[-140:]:
    print("%9.1f us  %s" % (us, name))
P
for kt in "" "4,1,16" "4,2,8" "8,1,8"; do
  echo "--- S2E_KTILE=$kt"
  S2E_KTILE=$kt timeout 300 python tools/conv_probe.py 32 81 49 256 512 4 32 41 25 256 512 4 32 161 97 256 128 2 32 321 193 64 64 2 32 81 49 512 256 2 2>&1 | tail -5
done > gpurun_out/c17_ktile_probe.log 2>&1
cat gpurun_out/c17_ktile_probe.log | cut -c1-200
