#!/bin/bash
# Round 2, first GPU call: the GPU gates on the shipped build, then the three-regime perf check of the shipped build
# and of the prepared single-flag variants (no per-variant gates in this pass: only a variant that wins gets gated).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "=== gates"; timeout -k 10 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== variants"; VARIANTS="${VARIANTS:-stack16 tex8 tri48 bvh4 bvh4_nosort spec_au_lv1 chunk_r20_b7}" bash scripts/perf_variants.sh
