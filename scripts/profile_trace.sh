#!/bin/bash
# ncu full capture of the closest-hit kernel only (4 launches: depths 0..3 of the second batch)
mkdir -p gpurun_out
WL=${1:-config2_mossy_ground}
TAG=${2:-r1}
CMD="python bench.py --workload $WL --spp 16 --steps 1 --warmup 3 --no-cpu --no-extra"
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 8 -c 4 -f -o gpurun_out/trace_${WL}_${TAG} $CMD > gpurun_out/trace_${WL}_${TAG}.log 2>&1
