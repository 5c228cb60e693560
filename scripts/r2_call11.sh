#!/bin/bash
# Round 2, eleventh GPU call: number of wavefronts (VOIDRAY_STREAMS=2..4) x persistent closest-hit blocks per SM
# (VOIDRAY_TRACE_BLOCKS=8,7,6: fewer leave registers for the other stream's kernels); bench line with kernel_alone.
mkdir -p gpurun_out
one() {  # one <workload> <spp> <steps>
  timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ka=(d['roofline'] or {}).get('kernel_alone') or {}
print('$1 spp $2: %.1f | %.1f  trace share %.3f  frac %.3f alone %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step'], d['roofline']['frac'] or 0, ka.get('frac') or 0))"
}
ab() { one config1_mushroom 64 10; one config2_mossy_ground 64 3; one config5_combined 16 2; }
for s in 2 3 4; do for b in 8 7 6; do
  echo "=== streams $s, trace blocks $b"; VOIDRAY_STREAMS=$s VOIDRAY_TRACE_BLOCKS=$b ab
done; done
echo "=== bench"; timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_call11.json 2> gpurun_out/bench_r2_call11.err; tail -c 600 gpurun_out/bench_r2_call11.err; cut -c1-300 gpurun_out/bench_r2_call11.json
