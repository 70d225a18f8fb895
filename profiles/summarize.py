"""Turn gpurun_out/ ncu artefacts into the small text summaries kept under profiles/.

  python profiles/summarize.py launches gpurun_out/t7_launches.csv  > profiles/r01_launches_<tag>.md
  python profiles/summarize.py full     gpurun_out/t7_raw.csv       > profiles/r01_ncu_full_<tag>.md

`launches` input: ncu --metrics gpu__time_duration.sum --csv log.  `full` input: ncu -i x.ncu-rep --page raw --csv.
Per-launch times under ncu are cold-cache and serialised: read SHARES, not absolutes.
"""
import csv
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__inst_executed_pipe_tma.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        if r[ui] == "us":
            v *= 1e3
        elif r[ui] == "ms":
            v *= 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot:.1f} % | {ns / 1e3 / n:.1f} |")
    print(f"\ntotal {tot / 1e3:.1f} us over {sum(a[0] for a in agg.values())} launches")


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"### {r[hdr.index('Kernel Name')]}\n")
        print("| metric | unit | value |\n|---|---|---|")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {units[i]} | {r[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
