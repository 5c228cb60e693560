#!/bin/bash
# Round 2, second GPU call: gates (scene-level culling, device group, RGBA8 textures as the shipped path), the new
# bench line (config 1 headline + the other configs), instruction counts for the issue-slot roofline, one full ncu
# capture of k_trace / k_shade on config 1, launch list of the bench command.
mkdir -p gpurun_out
echo "=== gates"; timeout -k 10 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== bench"; timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_call2.json 2> gpurun_out/bench_r2_call2.err; tail -c 600 gpurun_out/bench_r2_call2.err; cut -c1-1500 gpurun_out/bench_r2_call2.json
echo "=== reference arm"; timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
echo "=== instruction counts"; timeout -k 10 900 bash scripts/ncu_trace_inst.sh
echo "=== full capture, config 1"
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -c 6 -f -o gpurun_out/full_config1_r2 python scripts/render_once.py config1_mushroom 64 > gpurun_out/full_config1_r2.log 2>&1
echo "=== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config1_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/launches_config1_r2.log 2>&1
tail -c 300 gpurun_out/launches_config1_r2.log
ls -la gpurun_out | tail -20
