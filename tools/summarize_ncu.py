"""Turn ncu outputs into the small text summaries kept under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/<name>_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep > profiles/<name>_full.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    tot = sum(ns for _, ns in rows)
    agg = OrderedDict()
    for k, ns in rows:
        name = k.split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (name, cnt, ns / 1e6, 100 * ns / tot))
    print("\n%d launches, %.3f ms of kernel time in the captured window" % (len(rows), tot / 1e6))


WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "dram__sectors_read.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def full(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("### %s  (id %s)\n" % (d.get("Kernel Name", "?"), d.get("ID", "?")))
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for i, h in enumerate(hdr):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                print("| %s | %s | %s |" % (h, r[i], units[i]))
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
