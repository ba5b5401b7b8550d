"""Turns ncu CSV exports (gpurun_out/) into the small summaries committed under profiles/."""
import collections
import csv
import sys


def launch_summary(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        t = t / 1000 if unit == "ns" else (t * 1000 if unit == "ms" else t)
        name = row["Kernel Name"].split("(")[0].replace("void ", "")[:70]
        agg[name][0] += 1
        agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {path}; total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
        f.write("kernel,launches,total_us,avg_us,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{v[0]},{v[1]:.1f},{v[1] / v[0]:.2f},{100 * v[1] / tot:.1f}\n")


WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def raw_summary(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none; selected counters per profiled launch\n")
        f.write(",".join(f"{c} [{units[idx[c]]}]" for c in cols) + "\n")
        for r in rows[2:]:
            f.write(",".join('"' + r[idx[c]].replace('"', "'") + '"' for c in cols) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launch_summary(sys.argv[2], sys.argv[3])
    else:
        raw_summary(sys.argv[2], sys.argv[3])
