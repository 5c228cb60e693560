#!/bin/bash
# Round 2, nineteenth GPU call: triangle records bypass the L1 (no_allocate) with the surface id taken from the registers
# instead of a re-read — with (l1hints2) and without (l1tri) evict_last on the nodes, and with the ray loads bypassing
# the L1 as well (l1rays) — against the default build; two runs each. Gates on the variants first.
mkdir -p gpurun_out
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ka=(d['roofline'] or {}).get('kernel_alone') or {}
print('$1 spp $2: %.1f | %.1f  trace share %.3f  frac %.3f alone %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step'], d['roofline']['frac'] or 0, ka.get('frac') or 0))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config3_materials 64 3; one config5_combined 16 3; one config4_field 16 3; }
echo "=== gates (closest hit, variants)"; for v in l1tri l1hints2 l1rays; do VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/$v.so timeout -k 10 600 python -m pytest tests/test_gpu_closest_hit.py -q -m gpu -x 2>&1 | tail -1; done
for rep in 1 2; do
echo "=== default build"; ab
for v in l1tri l1hints2 l1rays; do echo "=== $v"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/$v.so ab; done
done
