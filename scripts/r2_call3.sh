#!/bin/bash
# Round 2, third GPU call: gates on the queue-order ray layout + tail kernel + device group; tail threshold sweep;
# bench line; instruction counts of every config for the issue-slot roofline.
mkdir -p gpurun_out
echo "=== gates"; timeout -k 10 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== tail threshold sweep (Msamples/s device | e2e)"
for cfg in "config1_mushroom 64 10" "config2_mossy_ground 64 3" "config5_combined 16 2"; do
  set -- $cfg
  for tm in 0 65536 262144 1048576 4194304; do
    VOIDRAY_TAIL_MAX=$tm timeout -k 10 300 python bench.py --workload $1 --spp $2 --steps $3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 spp $2 tail_max $tm: %.1f | %.1f  trace share %.3f' % (d['value'], d['e2e']['value'], d['roofline']['trace_share_of_step']))"
  done
done
echo "=== bench"; timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_call3.json 2> gpurun_out/bench_r2_call3.err; tail -c 600 gpurun_out/bench_r2_call3.err; cut -c1-700 gpurun_out/bench_r2_call3.json
echo "=== instruction counts"; timeout -k 10 1200 bash scripts/ncu_trace_inst.sh
echo "=== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config1_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/launches_config1_r2.log 2>&1
ls -la gpurun_out | tail -25
