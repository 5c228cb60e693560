#!/bin/bash
# On the GPU box: scripts/perf_check.sh for the default build and every library under gpurun_variants/
# (built here by scripts/build_variants.sh). PARITY=1 first runs the closest-hit and radiance gates on each variant,
# so that a variant that is faster but wrong is seen as such.
#   gpurun --timeout 900 -- 'PARITY=1 bash scripts/perf_variants.sh > gpurun_out/variants.log 2>&1'
#   VARIANTS="bvh4 bvh4_steps1" ... restricts the run to the named variants (default: every library built)
libs="default"
if [ -n "$VARIANTS" ]; then for v in $VARIANTS; do libs="$libs gpurun_variants/$v.so"; done; else libs="default $(ls gpurun_variants/*.so)"; fi
for lib in $libs; do
  if [ "$lib" = default ]; then unset VOIDRAY_CUDA_LIB; else export VOIDRAY_CUDA_LIB=$PWD/$lib; fi
  echo "== $lib"
  if [ -n "$PARITY" ] && [ "$lib" != default ]; then
    # a variant that hangs must not take the box with it: bounded, and its perf check is skipped if the gates fail
    timeout -k 10 600 python -m pytest tests/test_gpu_closest_hit.py tests/test_gpu_radiance.py -x -q -m gpu 2>&1 | tail -2
    if [ "${PIPESTATUS[0]}" != 0 ]; then echo "gates failed or timed out: perf check skipped"; continue; fi
  fi
  timeout -k 10 600 bash scripts/perf_check.sh
done
