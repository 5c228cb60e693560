#!/bin/bash
# Round 2, tenth GPU call: full gates on the two-wavefront build, bench line, reference arm, instruction counts of the
# shipped kernels on every config, launch list of the bench command.
mkdir -p gpurun_out
echo "=== gates"; timeout -k 10 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "=== bench"; timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_call10.json 2> gpurun_out/bench_r2_call10.err; tail -c 600 gpurun_out/bench_r2_call10.err; cut -c1-300 gpurun_out/bench_r2_call10.json
echo "=== reference arm"; timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-300
echo "=== instruction counts"; timeout -k 10 1200 bash scripts/ncu_trace_inst.sh
echo "=== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config1_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/launches_config1_r2.log 2>&1
tail -c 200 gpurun_out/launches_config1_r2.log
