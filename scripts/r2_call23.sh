#!/bin/bash
# Round 2, twenty-third GPU call: L2 eviction priority evict_last on the node and triangle-record fetches of k_trace
# (256-bit loads take the qualifier without a policy register) against the previous commit; two runs each.
mkdir -p gpurun_out
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ka=(d['roofline'] or {}).get('kernel_alone') or {}
print('$1 spp $2: %.1f | %.1f  trace share %.3f  frac %.3f alone %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step'], d['roofline']['frac'] or 0, ka.get('frac') or 0))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config5_combined 16 3; one config4_field 16 3; }
echo "=== gates (closest hit)"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/l2keep.so timeout -k 10 600 python -m pytest tests/test_gpu_closest_hit.py -q -m gpu -x 2>&1 | tail -1
for rep in 1 2; do
echo "=== base"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/base.so ab
echo "=== l2keep"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/l2keep.so ab
done
