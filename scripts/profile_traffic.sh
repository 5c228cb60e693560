#!/bin/bash
# One ncu --set full capture of all 8 k_trace launches of one wavefront batch (depths 0..7) -> DRAM traffic per launch.
mkdir -p gpurun_out
WL=${1:-config2_mossy_ground}
TAG=${2:-r1}
SPP=${3:-16}
CMD="python bench.py --workload $WL --spp $SPP --steps 1 --warmup 3 --no-cpu --no-extra"
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 8 -c 8 -f -o gpurun_out/traffic_${WL}_${TAG} $CMD > gpurun_out/traffic_${WL}_${TAG}.log 2>&1
tail -2 gpurun_out/traffic_${WL}_${TAG}.log
