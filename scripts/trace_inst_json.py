"""gpurun_out/inst_<cfg>.{csv,json} (scripts/ncu_trace_inst.sh) -> profiles/r2_trace_inst.json:
per config, warp / thread instructions and DRAM bytes of k_trace per ray segment (all depths of one accumulate), plus the
same per launch as a table in profiles/r2_trace_inst.md."""
import collections
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
out, md = {}, ["# k_trace / k_shade instruction and DRAM counts per launch (ncu, --clock-control none)\n"]
for jp in sorted(glob.glob(os.path.join(src, "inst_*.json"))):
    name = os.path.basename(jp)[5:-5]
    try:
        meta = json.loads(open(jp).read().strip().splitlines()[-1])
    except Exception:
        continue
    rows = [r for r in csv.reader(open(jp[:-5] + ".csv")) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    idx = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in data:
        key = (int(r[idx["ID"]]), r[idx["Kernel Name"]].split("(")[0].replace("void ", ""))
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        if unit in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        if unit in ("us", "ms", "ns") and r[idx["Metric Name"]].startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}[unit]  # -> us
        launches.setdefault(key, {})[r[idx["Metric Name"]]] = v
    tot = collections.defaultdict(float)
    md.append(f"\n## {name}: {meta['width']}x{meta['height']} x {meta['spp']} spp, {meta['segments_per_repeat']} segments\n")
    md.append("| launch | kernel | us | warp inst | lanes / inst | issue busy % | DRAM MB |\n|---:|---|---:|---:|---:|---:|---:|")
    for (lid, kn), m in launches.items():
        wi, ti = m.get("smsp__inst_executed.sum", 0.0), m.get("smsp__thread_inst_executed.sum", 0.0)
        dram = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        md.append(f"| {lid} | {kn} | {m.get('gpu__time_duration.sum', 0):.1f} | {wi:.0f} | {ti / wi if wi else 0:.2f} | "
                  f"{m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} | {dram / 1e6:.1f} |")
        k = "trace" if "k_trace" in kn else ("shade" if ("k_shade" in kn or "k_miss" in kn) else "other")
        tot[k + "_warp"] += wi
        tot[k + "_thread"] += ti
        tot[k + "_dram"] += dram
        tot[k + "_us"] += m.get("gpu__time_duration.sum", 0.0)
    seg = float(meta["segments_per_repeat"]) * meta.get("repeats", 1)
    out[name] = {"warp_inst_per_segment": tot["trace_warp"] / seg, "thread_inst_per_segment": tot["trace_thread"] / seg,
                 "dram_bytes_per_segment": tot["trace_dram"] / seg,
                 "shade_warp_inst_per_segment": tot["shade_warp"] / seg, "shade_thread_inst_per_segment": tot["shade_thread"] / seg,
                 "shade_dram_bytes_per_segment": tot["shade_dram"] / seg,
                 "trace_us_under_ncu": tot["trace_us"], "shade_us_under_ncu": tot["shade_us"],
                 "other_us_under_ncu": tot["other_us"], "segments": seg,
                 "source": f"ncu smsp__inst_executed.sum / smsp__thread_inst_executed.sum / dram__bytes over every k_trace launch of "
                           f"one accumulate, {meta['width']}x{meta['height']} x {meta['spp']} spp (scripts/ncu_trace_inst.sh)"}
    md.append(f"\nper segment: k_trace {out[name]['warp_inst_per_segment']:.1f} warp instructions "
              f"({out[name]['thread_inst_per_segment'] / 32 / out[name]['warp_inst_per_segment']:.3f} lane efficiency), "
              f"{out[name]['dram_bytes_per_segment']:.1f} DRAM bytes; k_shade + k_miss {out[name]['shade_warp_inst_per_segment']:.1f} warp "
              f"instructions, {out[name]['shade_dram_bytes_per_segment']:.1f} DRAM bytes")
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_trace_inst.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", "r2_trace_inst.md"), "w").write("\n".join(md) + "\n")
print(json.dumps(out, indent=1))
