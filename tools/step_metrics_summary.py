"""Summarise `ncu --metrics <list> --csv` of one eager training step: one line per kernel class (kernel name x grid size) with
its launch count and total time, and the metrics of its LONGEST launch (DRAM bytes, DRAM %, tensor-pipe %, registers, ...).

  python tools/step_metrics_summary.py gpurun_out/step_metrics.csv > profiles/rXX_step_metrics_summary.txt
"""
import collections
import csv
import re
import sys

rows = csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("=="))
launch = collections.OrderedDict()      # ID -> {name, grid, metrics}
for r in rows:
    e = launch.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", ""), "m": {}})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)      # -> us
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)                  # -> MB
    e["m"][r["Metric Name"]] = v


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("<unnamed>::", "")
    return n[:58]


groups = collections.OrderedDict()
for e in launch.values():
    groups.setdefault((short(e["name"]), e["grid"], e["block"]), []).append(e["m"])
T = sum(m.get("gpu__time_duration.sum", 0.0) for g in groups.values() for m in g)
print("# %d launches, %.2f ms serialised (ncu, cold caches: compare shares and per-launch metrics, not the sum)" % (len(launch), T / 1e3))
print("# per class: launches, total us, share | longest launch: us, DRAM read MB, write MB, DRAM %, tensor pipe %, warps active %, L2 hit %, regs")
order = sorted(groups.items(), key=lambda kv: -sum(m.get("gpu__time_duration.sum", 0.0) for m in kv[1]))
for (name, grid, block), ms in order:
    tot = sum(m.get("gpu__time_duration.sum", 0.0) for m in ms)
    if tot < 0.0005 * T:
        continue
    b = max(ms, key=lambda m: m.get("gpu__time_duration.sum", 0.0))
    g = lambda k: b.get(k, float("nan"))
    print("%-58s grid %-12s x%-3d %9.0f us %5.1f%% | %8.1f us  rd %8.1f  wr %8.1f  dram %5.1f%%  tensor %5.1f%%  warps %5.1f%%  L2hit %5.1f%%  regs %3.0f" % (
        name, grid.replace(" ", ""), len(ms), tot, 100 * tot / T, g("gpu__time_duration.sum"), g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
        g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        g("sm__warps_active.avg.pct_of_peak_sustained_active"), g("lts__t_sector_hit_rate.pct"), g("launch__registers_per_thread")))
