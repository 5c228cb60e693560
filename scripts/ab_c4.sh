#!/bin/bash
for lib in default gpurun_variants/lib_prefma.so; do
  if [ "$lib" = default ]; then unset VOIDRAY_CUDA_LIB; else export VOIDRAY_CUDA_LIB=$PWD/$lib; fi
  for spp in 16 64; do
  python bench.py --workload config4_field --spp $spp --steps 2 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib spp $spp', round(d['value'],1), 'Msamples/s', round(d['mrays_per_s'],1), 'Mrays/s ms/step', round(d['ms_per_step'],1), 'commit', d['e2e']['commit_ms'])"
  done
done
