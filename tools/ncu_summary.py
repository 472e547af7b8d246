"""Condense `ncu --page raw --csv` exports (one row per profiled launch, ~1500 metric columns) into the handful of numbers
the roofline discussion uses.   python tools/ncu_summary.py gpurun_out/r02_ncu_x.csv [...] > profiles/r02_ncu_x.txt"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("sm__cycles_elapsed.max.per_second", "sm_GHz"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier")]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print("# %s" % path)
        for r in rows[2:]:
            print(r[idx["Kernel Name"]].replace("void <unnamed>::", "")[:110])
            print("    " + "  ".join("%s=%s%s" % (lab, r[idx[k]], (" " + units[idx[k]]) if units[idx[k]] not in ("", "%") else "")
                                      for k, lab in KEYS if k in idx))
        print()


if __name__ == "__main__":
    main()
