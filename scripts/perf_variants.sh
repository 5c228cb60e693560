#!/bin/bash
for lib in default gpurun_variants/*.so; do
  if [ "$lib" = default ]; then unset VOIDRAY_CUDA_LIB; else export VOIDRAY_CUDA_LIB=$PWD/$lib; fi
  echo "== $lib"; bash scripts/perf_check.sh
done
