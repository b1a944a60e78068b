#!/usr/bin/env python3
"""Summarise an ncu capture for profiles/: key raw metrics of the top kernel
(from <name>.ncu-rep) and per-kernel time shares (from the launch list CSV).
usage: tools/ncu_summary.py <rep> <launches.csv> <out.md> [title]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
]


def main():
    rep, launches, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    lines = ["# " + title, ""]
    if len(rows) >= 3:
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            lines.append("## kernel `%s`" % d.get("Kernel Name", "?")[:100])
            lines.append("")
            lines.append("| metric | value | unit |")
            lines.append("|---|---|---|")
            u = dict(zip(hdr, units))
            for k in KEYS:
                for h in hdr:
                    if h == k or h.endswith("." + k):
                        lines.append("| %s | %s | %s |" % (k, d[h], u.get(h, "")))
                        break
            lines.append("")
    # launch list
    try:
        txt = open(launches).read()
        body = txt[txt.index('"ID"'):]
        rd = csv.DictReader(io.StringIO(body))
        tot = defaultdict(float)
        cnt = defaultdict(int)
        for r in rd:
            if r.get("Metric Name") == "gpu__time_duration.sum":
                v = float(r["Metric Value"].replace(",", ""))
                unit = r.get("Metric Unit", "ns")
                scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
                name = r["Kernel Name"].split("(")[0][:70]
                tot[name] += v * scale
                cnt[name] += 1
        all_ms = sum(tot.values())
        lines += ["## launch list (gpu__time_duration.sum, --clock-control none; shares, not absolutes)", "",
                  "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            lines.append("| %s | %d | %.3f | %.1f%% |" % (k, cnt[k], v, 100 * v / all_ms if all_ms else 0))
    except (OSError, ValueError) as e:
        lines.append("launch list unavailable: %s" % e)
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
