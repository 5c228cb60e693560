#!/bin/bash
# bench each kernel variant in gpurun_variants/ (plus the default build)
for lib in default gpurun_variants/*.so; do
  if [ "$lib" = default ]; then unset VOIDRAY_CUDA_LIB; else export VOIDRAY_CUDA_LIB=$PWD/$lib; fi
  python bench.py --spp 64 --steps 3 --no-cpu --no-extra "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', round(d['value'],1), 'Msamples/s', round(d['mrays_per_s'],1), 'Mrays/s trace avg ms', round(d['roofline']['avg_launch_ms'],4), 'share', round(d['roofline']['trace_share_of_step'],3))"
done
