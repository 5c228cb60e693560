#!/bin/bash
# Round 2, thirteenth GPU call: k_trace stacks sized by the tree depth (more of the SM's 256 KB stays L1), the 30-degree
# decision without atan2 outside a guard band, L2 set-aside for the HDRI during k_miss (VOIDRAY_L2_PERSIST=<MB>).
mkdir -p gpurun_out
echo "=== gates"; timeout -k 10 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ka=(d['roofline'] or {}).get('kernel_alone') or {}
print('$1 spp $2: %.1f | %.1f  trace share %.3f  frac %.3f alone %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step'], d['roofline']['frac'] or 0, ka.get('frac') or 0))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config3_materials 64 3; one config5_combined 16 2; one config4_field 16 3; }
echo "=== base (previous commit)"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/base.so ab
echo "=== new"; ab
echo "=== new + L2 persist 48 MB"; VOIDRAY_L2_PERSIST=48 ab
echo "=== base again"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/base.so ab
echo "=== new again"; ab
