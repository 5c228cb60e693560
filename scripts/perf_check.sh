#!/bin/bash
# quick perf check of the three regimes: small scene (config1), L2-resident textured (config2), 10M triangles (config4)
for cfg in "config1_mushroom 64" "config2_mossy_ground 64" "config4_field 16"; do
  set -- $cfg
  python bench.py --workload $1 --spp $2 --steps 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline'] or {}
print('$1 spp $2:', round(d['value'],1), 'Msamples/s', round(d['mrays_per_s'],1), 'Mrays/s  trace avg ms', round(r.get('avg_launch_ms',0),4), 'share', round(r.get('trace_share_of_step',0),3), 'e2e', round(d['e2e']['value'],1))"
done
