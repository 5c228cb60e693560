#!/bin/bash
# Round 2, seventh GPU call: k_shade_first at depth 0 + slot-major attenuation stack; refill-threshold sweep per depth
# class (VOIDRAY_REFILL="<depth 0>,<deeper>"); instruction counts of the default build.
mkdir -p gpurun_out
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 spp $2: %.1f | %.1f  trace share %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step']))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config5_combined 16 2; }
echo "=== gates"; timeout -k 10 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== base"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/base.so ab
echo "=== new (default build)"; ab; one config3_materials 64 3; one config4_field 16 3
for r in 16,12 20,12 24,12 12,8 12,16 12,20; do
  echo "=== refill $r"; VOIDRAY_REFILL=$r ab
done
echo "=== instruction counts"; CONFIGS="config1_mushroom:64 config2_mossy_ground:16" timeout -k 10 600 bash scripts/ncu_trace_inst.sh
