#!/bin/bash
# Round 2, sixth GPU call: k_shade with span-batched scan; A/B of the span / register-cap variants against the
# previous commit's library; instruction counts of the default build.
mkdir -p gpurun_out
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 spp $2: %.1f | %.1f  trace share %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step']))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config5_combined 16 2; }
echo "=== gates (radiance + api)"; timeout -k 10 900 python -m pytest tests/test_gpu_radiance.py tests/test_gpu_api.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -3
for v in ${VARIANTS:-base span2 span8 mb6}; do
  echo "=== $v"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/$v.so ab
done
echo "=== new (default build)"; ab; one config3_materials 64 3; one config4_field 16 3
echo "=== instruction counts"; CONFIGS="config1_mushroom:64 config2_mossy_ground:16" timeout -k 10 600 bash scripts/ncu_trace_inst.sh
