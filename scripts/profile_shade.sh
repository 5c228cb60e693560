#!/bin/bash
# ncu full capture (with source) of the shading kernel: depths 1..2 of the second batch
mkdir -p gpurun_out
WL=${1:-config2_mossy_ground}
TAG=${2:-r1}
CMD="python bench.py --workload $WL --spp 16 --steps 1 --warmup 3 --no-cpu --no-extra"
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 2 -f -o gpurun_out/shade_${WL}_${TAG} $CMD > gpurun_out/shade_${WL}_${TAG}.log 2>&1
