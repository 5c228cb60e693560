#!/bin/bash
# perf_check.sh for the default build and every library under gpurun_variants/
for lib in default gpurun_variants/*.so; do
  if [ "$lib" = default ]; then unset VOIDRAY_CUDA_LIB; else export VOIDRAY_CUDA_LIB=$PWD/$lib; fi
  echo "== $lib"; bash scripts/perf_check.sh
done
