#!/bin/bash
# bench.py over all five BASELINE configs at their own spp (N = 1): one JSON line each -> gpurun_out/bench_all_<tag>.jsonl
TAG=${1:-r1}
OUT=gpurun_out/bench_all_${TAG}.jsonl
mkdir -p gpurun_out; : > $OUT
python bench.py --workload config1_mushroom --steps 5 --no-extra 2>/dev/null | tail -1 >> $OUT
python bench.py --workload config2_mossy_ground --steps 3 --no-extra 2>/dev/null | tail -1 >> $OUT
python bench.py --workload config3_materials --steps 2 --no-extra 2>/dev/null | tail -1 >> $OUT
python bench.py --workload config4_field --steps 2 --no-extra 2>/dev/null | tail -1 >> $OUT
python bench.py --workload config5_combined --steps 2 --no-extra 2>/dev/null | tail -1 >> $OUT
TAG=$TAG python - <<'P'
import json, os
for line in open("gpurun_out/bench_all_%s.jsonl" % os.environ["TAG"]):
    d = json.loads(line)
    r = d.get("roofline") or {}
    c = d.get("cpu_baseline") or {}
    print(d["config"]["workload"], "spp", d["config"]["spp_per_gpu"], "| %.0f Msamples/s %.0f Mrays/s | e2e %.0f | roofline frac %s | cpu port %s (%s cores)" % (
        d["value"], d["mrays_per_s"], d["e2e"]["value"], ("%.2f" % r["frac"]) if r else "n/a", ("%.2f" % c["value"]) if c else "n/a", c.get("cores")))
P
