#!/bin/bash
for cfg in "4 1.0" "4 0.5" "4 2.0" "2 1.0" "1 1.0" "7 1.0" "7 2.0" "7 3.0"; do
  set -- $cfg
  export VOIDRAY_LEAF_MAX=$1 VOIDRAY_NODE_COST=$2
  python bench.py --spp 64 --steps 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('leaf_max $1 node_cost $2', round(d['value'],1), 'Msamples/s trace avg ms', round(d['roofline']['avg_launch_ms'],4), 'nodes', d['config']['bvh_nodes'])"
done
