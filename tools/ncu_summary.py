"""Summarise ncu outputs into small text files for profiles/ (run in the build container, no GPU).

    python tools/ncu_summary.py launches <launches.csv>            # per-kernel share of the step
    python tools/ncu_summary.py raw <report.ncu-rep> [name filter]  # key metrics per captured launch
    python tools/ncu_summary.py sass <report.ncu-rep>               # executed warp instructions by opcode
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "smsp__issue_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {v[0]:8d} {v[1]:12.1f} {v[1] / tot:7.1%} {v[1] / v[0]:9.1f}")
    print(f"{'TOTAL':60s} {sum(v[0] for v in agg.values()):8d} {tot:12.1f}")


def raw(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    print("kernel launches:", [r[ci["Kernel Name"]].split("(")[0] for r in rows[2:]])
    for k in KEYS:
        if k in ci:
            print(f"{k:82s} {units[ci[k]]:16s} {[r[ci[k]] for r in rows[2:]]}")


def sass(rep):
    rows = ncu_csv(rep, "source")
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            kern.append(cur)
        elif cur is not None:
            cur.append(r)
    k = kern[0]
    ci = {h: i for i, h in enumerate(k[0])}
    ops, smp = collections.Counter(), collections.Counter()
    for r in k[1:]:
        if len(r) < len(k[0]):
            continue
        src = r[ci["Source"]].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        m = src.split()[0].split(".")[0]
        ops[m] += int(r[ci["Instructions Executed"]] or 0)
        smp[m] += int(r[ci["# Samples"]] or 0)
    tot, stot = sum(ops.values()), max(1, sum(smp.values()))
    print(f"first captured launch: {tot} warp instructions, {stot} stall samples")
    for m, n in ops.most_common(20):
        print(f"{m:10s} {n:12d} {n / tot:6.1%} of instructions  {smp[m] / stot:6.1%} of samples")


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "sass": sass}[sys.argv[1]](sys.argv[2])
