"""Summarise ncu captures under gpurun_out/ into tracked files under profiles/.
  python scripts/summarize_profile.py launches <launches.csv> <out.md>
  python scripts/summarize_profile.py kernel <raw.csv (ncu --page raw --csv)> <out.md> [title]"""
import collections
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads active per instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    seq = []
    for r in data:
        n = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(n, [0, 0.0])
        agg[n][0] += 1
        agg[n][1] += v
        seq.append((n, v))
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"ncu launch list `{path}` ({len(seq)} launches, gpu__time_duration.sum, --clock-control none).\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {n} | {c} | {t:.1f} | {t / tot:.3f} |\n")
        f.write("\nFirst wavefront batch, launch by launch (us):\n\n```\n")
        for n, v in seq[:22]:
            f.write(f"{n:16s} {v:9.1f}\n")
        f.write("```\n")


def kernel(path, out, title):
    rows = list(csv.reader(open(path)))
    h, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"{title}\n\nSource: `ncu --set full --clock-control none --import-source on`, raw page `{path}`; "
                f"one column per captured launch.\n\n")
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---:|" * len(data) + "\n")
        if "Kernel Name" in h:
            f.write("| kernel | | " + " | ".join(d[h.index("Kernel Name")].split("(")[0].replace("void ", "") for d in data) + " |\n")
        for key, label in KEYS:
            if key in h:
                i = h.index(key)
                f.write(f"| {label} (`{key}`) | {units[i]} | " + " | ".join(d[i] for d in data) + " |\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu capture")
