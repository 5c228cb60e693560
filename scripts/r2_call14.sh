#!/bin/bash
# Round 2, fourteenth GPU call: the 30-degree decision without atan2 outside a guard band (default build) and the
# one-instruction node address (nodeaddr.so) against the previous commit (base.so), two runs each.
mkdir -p gpurun_out
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ka=(d['roofline'] or {}).get('kernel_alone') or {}
print('$1 spp $2: %.1f | %.1f  trace share %.3f  frac %.3f alone %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step'], d['roofline']['frac'] or 0, ka.get('frac') or 0))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config3_materials 64 3; one config5_combined 16 2; one config4_field 16 3; }
echo "=== gates (radiance)"; timeout -k 10 900 python -m pytest tests/test_gpu_radiance.py tests/test_gpu_full_size.py tests/test_gpu_units.py -q -m gpu -x 2>&1 | tail -3
for rep in 1 2; do
echo "=== base (previous commit)"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/base.so ab
echo "=== angle shortcut (default build)"; ab
echo "=== angle shortcut + node address"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/nodeaddr.so ab
done
