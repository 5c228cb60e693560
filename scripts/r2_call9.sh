#!/bin/bash
# Round 2, ninth GPU call: two wavefronts on two streams (batches alternate, accumulation in batch order) against
# VOIDRAY_STREAMS=1 (single wavefront); full gates on the default (dual) build.
mkdir -p gpurun_out
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 spp $2: %.1f | %.1f  trace share %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step']))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config3_materials 64 3; one config5_combined 16 2; one config4_field 16 3; }
echo "=== gates"; timeout -k 10 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== single wavefront"; VOIDRAY_STREAMS=1 ab
echo "=== two wavefronts (default)"; ab
echo "=== single wavefront again"; VOIDRAY_STREAMS=1 ab
echo "=== two wavefronts again"; ab
echo "=== bench"; timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_call9.json 2> gpurun_out/bench_r2_call9.err; tail -c 600 gpurun_out/bench_r2_call9.err; cut -c1-400 gpurun_out/bench_r2_call9.json
