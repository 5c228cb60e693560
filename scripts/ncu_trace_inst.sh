#!/bin/bash
# On the GPU box: warp / thread instruction counts and DRAM bytes of every k_trace (and k_shade) launch of one accumulate,
# per BASELINE config -> gpurun_out/inst_<cfg>.csv + .json; scripts/trace_inst_json.py folds them into
# profiles/r2_trace_inst.json (what bench.py's issue-slot roofline reads).
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max
CONFIGS=${CONFIGS:-"config1_mushroom:64 config2_mossy_ground:16 config3_materials:16 config5_combined:4 config4_field:8"}
for cfg in $CONFIGS; do
  name=${cfg%%:*}; spp=${cfg##*:}
  ncu --metrics $M --clock-control none -k regex:'k_trace|k_shade|k_miss|k_raygen|k_accumulate' --csv --log-file gpurun_out/inst_${name}.csv \
      python scripts/render_once.py $name $spp > gpurun_out/inst_${name}.json 2> gpurun_out/inst_${name}.err
  tail -1 gpurun_out/inst_${name}.json | cut -c1-200
done
