#!/bin/bash
for paths in 4194304 8388608 16777216 33554432 67108864; do
  for cfg in "config1_mushroom 64" "config2_mossy_ground 64"; do
  set -- $cfg
  python bench.py --workload $1 --spp $2 --steps 3 --no-cpu --no-extra --paths $paths 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('paths $paths $1:', round(d['value'],1), 'Msamples/s', round(d['mrays_per_s'],1), 'Mrays/s launches', d['gpu_launches'])"
  done
done
