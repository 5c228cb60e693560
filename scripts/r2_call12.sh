#!/bin/bash
# Round 2, twelfth GPU call: full-size config 4 parity as a pytest (timed), compute-sanitizer over smoke(), k_trace block
# size 64 / 256 against the shipped 128.
mkdir -p gpurun_out
echo "=== config 4 at full size"; timeout -k 10 1500 python -m pytest tests/test_gpu_config4_full.py -q -m gpu --durations=3 2>&1 | tail -8
echo "=== sanitizer"; bash scripts/sanitize.sh
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ka=(d['roofline'] or {}).get('kernel_alone') or {}
print('$1 spp $2: %.1f | %.1f  trace share %.3f  frac %.3f alone %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step'], d['roofline']['frac'] or 0, ka.get('frac') or 0))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config5_combined 16 2; }
echo "=== default (128 threads x 8 blocks)"; ab
echo "=== 64 threads x 16 blocks"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/trace64.so ab
echo "=== 256 threads x 4 blocks"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/trace256.so ab
