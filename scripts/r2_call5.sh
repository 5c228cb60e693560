#!/bin/bash
# Round 2, fifth GPU call: gates on the new k_shade (miss list + per-warp hit compaction), k_raygen (one pixel per
# thread), fast division, and the node step without the widening multiply; A/B against the previous commit's library
# (gpurun_variants/base.so); instruction counts; launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "=== gates"; timeout -k 10 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
ab() {
  for cfg in "config1_mushroom 64 10" "config2_mossy_ground 64 3" "config3_materials 64 3" "config4_field 16 3" "config5_combined 16 2"; do
    set -- $cfg
    timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 spp $2: %.1f | %.1f  trace share %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step']))"
  done
}
echo "=== base (previous commit)"; VOIDRAY_CUDA_LIB=$PWD/gpurun_variants/base.so ab
echo "=== new"; ab
echo "=== bench"; timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_call5.json 2> gpurun_out/bench_r2_call5.err; tail -c 600 gpurun_out/bench_r2_call5.err; cut -c1-700 gpurun_out/bench_r2_call5.json
echo "=== instruction counts"; CONFIGS="config1_mushroom:64 config2_mossy_ground:16 config3_materials:16 config5_combined:4 config4_field:8" timeout -k 10 1200 bash scripts/ncu_trace_inst.sh
ls -la gpurun_out | tail -25
