#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list + one full capture of the closest-hit kernel.
# Outputs land in gpurun_out/; summaries are copied to profiles/ by scripts/summarize_profile.py.
set -x
mkdir -p gpurun_out
WL=${1:-config2_mossy_ground}
TAG=${2:-r1}
CMD="python bench.py --workload $WL --spp 16 --steps 1 --warmup 3 --no-cpu --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv $CMD > gpurun_out/launches_${WL}_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 8 -c 4 -f -o gpurun_out/trace_${WL}_${TAG} $CMD > gpurun_out/trace_${WL}_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 2 -f -o gpurun_out/shade_${WL}_${TAG} $CMD > gpurun_out/shade_${WL}_${TAG}.log 2>&1
ls -la gpurun_out
