#!/usr/bin/env python3
"""Turn ncu CSV exports into the markdown tables committed under profiles/.

    python profiles/summarize_ncu.py launches  gpurun_out/x.csv  > profiles/rNN_ncu_launches_*.md
    python profiles/summarize_ncu.py full      gpurun_out/x_raw.csv > profiles/rNN_ncu_full_*.md

`launches`: the `--metrics gpu__time_duration.sum --clock-control none` pass (one row per launch);
`full`: `ncu -i rep --page raw --csv` of an `ncu --set full` capture."""
import collections
import csv
import io
import sys


def rows_of(path):
    txt = open(path).read().splitlines()
    start = next(i for i, l in enumerate(txt) if l.startswith('"ID"'))
    return list(csv.reader(io.StringIO("\n".join(txt[start:]))))


def launches(path):
    rows = rows_of(path)
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    by = collections.OrderedDict()
    for r in rows[1:]:
        by.setdefault(r[k], []).append(float(r[v]))
    total = sum(sum(x) for x in by.values())
    print("| kernel | launches | mean us (cold cache, serialised) | share of all launches |")
    print("|---|---|---|---|")
    for name, x in by.items():
        print(f"| `{name[:150]}` | {len(x)} | {sum(x) / len(x) / 1e3:.1f} | {sum(x) / total * 100:.2f} % |")
    print(f"\n{sum(len(x) for x in by.values())} launches, {total / 1e6:.3f} ms in total")


FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def full(path):
    rows = rows_of(path)
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(m, hdr.index(m)) for m in FULL if m in hdr]
    k = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join(f"{m} [{units[i]}]" for m, i in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in data:
        print(f"| `{r[k][:110]}` | " + " | ".join(r[i] for _m, i in cols) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
