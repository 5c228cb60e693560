#!/bin/bash
# Round 2, twenty-fifth GPU call: camera-ray culling in k_raygen, switched by the host on the culled fraction of the
# previous call (off below 30 %, probe every 16th call): full gates, then A/B against the previous commit.
mkdir -p gpurun_out
echo "=== gates"; timeout -k 10 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ka=(d['roofline'] or {}).get('kernel_alone') or {}
print('$1 spp $2: %.1f | %.1f  trace share %.3f  frac %.3f alone %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step'], d['roofline']['frac'] or 0, ka.get('frac') or 0))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 4; one config3_materials 64 4; one config5_combined 16 4; one config4_field 16 3; }
echo "=== base"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/base.so ab
echo "=== default build"; ab
