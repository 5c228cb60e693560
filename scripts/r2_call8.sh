#!/bin/bash
# Round 2, eighth GPU call: evict-first hints on the path-state streams (A/B), fast-integrator evaluation
# (time to equal error), full ncu capture of the shipped kernels on config 1 and config 2.
mkdir -p gpurun_out
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 spp $2: %.1f | %.1f  trace share %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step']))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config5_combined 16 2; one config4_field 16 3; }
echo "=== new (default build)"; ab
echo "=== stream hints"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/stream.so ab
echo "=== new (default build) again"; ab
echo "=== stream hints again"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/stream.so ab
echo "=== fast integrator, config 2 at 480x270"; timeout -k 10 600 python scripts/fast_integrator_eval.py config2_mossy_ground 480 270 > gpurun_out/fast_eval_config2.json 2> gpurun_out/fast_eval_config2.err; tail -c 300 gpurun_out/fast_eval_config2.err; cat gpurun_out/fast_eval_config2.json | tr -d '\n ' | cut -c1-1500; echo
echo "=== fast integrator, config 1 at 400x300"; timeout -k 10 600 python scripts/fast_integrator_eval.py config1_mushroom 400 300 > gpurun_out/fast_eval_config1.json 2> gpurun_out/fast_eval_config1.err; cat gpurun_out/fast_eval_config1.json | tr -d '\n ' | cut -c1-1500; echo
echo "=== full capture, config 1 (depth 0 and 1 of every kernel)"
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade|k_miss|k_raygen' -c 6 -f -o gpurun_out/full_config1_r2b python scripts/render_once.py config1_mushroom 64 > gpurun_out/full_config1_r2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_miss' -c 1 -f -o gpurun_out/miss_config2_r2b python scripts/render_once.py config2_mossy_ground 16 > gpurun_out/miss_config2_r2b.log 2>&1
ls -la gpurun_out | tail -12
