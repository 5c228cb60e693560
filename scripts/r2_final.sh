#!/bin/bash
# Round 2, final single-GPU call: full gates, smoke, bench line (+ reference arm), instruction counts of the shipped
# kernels on every config, ncu --set full of the shipped kernels on config 1 and of k_trace depth 0 / 1 on config 2,
# launch list of the bench command.
mkdir -p gpurun_out
echo "=== gates"; timeout -k 10 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "=== bench"; timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 600 gpurun_out/bench_r2_final.err; cut -c1-300 gpurun_out/bench_r2_final.json
echo "=== reference arm"; timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-300
echo "=== instruction counts"; timeout -k 10 1200 bash scripts/ncu_trace_inst.sh
echo "=== full capture, config 1"
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade|k_miss|k_raygen' -c 6 -f -o gpurun_out/full_config1_r2c python scripts/render_once.py config1_mushroom 64 > gpurun_out/full_config1_r2c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace' -c 2 -f -o gpurun_out/trace_config2_r2c python scripts/render_once.py config2_mossy_ground 16 > gpurun_out/trace_config2_r2c.log 2>&1
echo "=== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config1_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/launches_config1_r2.log 2>&1
tail -c 200 gpurun_out/launches_config1_r2.log
ls -la gpurun_out | tail -8
