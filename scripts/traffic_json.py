"""profiles/trace_traffic.json from an ncu --set full capture of the k_trace launches of one wavefront batch:
DRAM bytes (read + write) per launch, averaged over the captured launches (depths 0..7), next to the
algorithmic bytes of the same launches.  python scripts/traffic_json.py <workload> <file.ncu-rep>"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl, rep = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
def col(name):
    i = h.index(name)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units[i], 1.0)
    return [float(d[i].replace(",", "")) * scale for d in data]
rd, wr, dur = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
path = os.path.join(ROOT, "profiles", "trace_traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
per = [r + w for r, w in zip(rd, wr)]
out[wl] = sum(per) / len(per)
out[wl + "_detail"] = {"launches": len(per), "dram_bytes_per_launch": per, "duration_s": dur,
                       "dram_gbs_per_launch": [b / t / 1e9 for b, t in zip(per, dur)], "source": os.path.basename(rep)}
# the level the kernel actually works at when the tree is cache-resident: bytes L2 delivered to the L1s, hit rates
try:
    l2rd = col("l1tex__m_xbar2l1tex_read_bytes.sum")
    l1hit, l2hit = col("l1tex__t_sector_hit_rate.pct"), col("lts__t_sector_hit_rate.pct")
    out[wl + "_l2"] = {"l2_to_l1_read_bytes_per_launch": sum(l2rd) / len(l2rd),
                       "l2_to_l1_read_gbs": sum(l2rd) / sum(dur) / 1e9,
                       "l1_sector_hit_pct_per_launch": l1hit, "l2_sector_hit_pct_per_launch": l2hit,
                       "source": os.path.basename(rep)}
except ValueError:
    pass
json.dump(out, open(path, "w"), indent=1)
print(wl, "avg DRAM bytes/launch %.1f MB" % (out[wl] / 1e6), "GB/s per launch", [round(x) for x in out[wl + "_detail"]["dram_gbs_per_launch"]])
