#!/usr/bin/env python3
"""One markdown table row per kernel launch of an ncu capture (all launches of a .ncu-rep): time, grid, registers,
issue / pipe / memory utilisation and DRAM bytes.  usage: tools/ncu_table.py <rep> [<rep> ...] >> profiles/<name>.md"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM written"), ("smsp__inst_executed.sum", "warp instructions")]


def main():
    print("| kernel | " + " | ".join(t for _, t in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        head, units = rows[0], rows[1]
        ki = head.index("Kernel Name")
        for r in rows[2:]:
            cells = []
            for key, _ in COLS:
                if key in head:
                    i = head.index(key)
                    v = r[i]
                    try:
                        v = "%.4g" % float(v.replace(",", ""))
                    except ValueError:
                        pass
                    cells.append("%s %s" % (v, units[i]) if units[i] not in ("", "%") and (key.startswith("dram__bytes") or key.startswith("gpu__time")) else v)
                else:
                    cells.append("")
            name = r[ki].split("(")[0].replace("unnamed>::", "").replace("void ", "")
            print("| `%s` | " % name + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
