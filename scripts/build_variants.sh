#!/bin/bash
# Builds experiment variants of libvoidray_cuda.so into gpurun_variants/ (travels to the GPU box with the snapshot;
# scripts/perf_variants.sh then runs scripts/perf_check.sh on the default build and on each of them through
# VOIDRAY_CUDA_LIB). Each variant is one set of -D flags applied to every translation unit.
#   bash scripts/build_variants.sh                # all variants below
#   bash scripts/build_variants.sh stack16        # only the named ones
set -e
cd "$(dirname "$0")/.."
declare -A FLAGS=(
  [stack16]="-DVR_SMEM_STACK=16"
  [stack12]="-DVR_SMEM_STACK=12"
  [tri48]="-DVR_TRI48"
  [stack16_tri48]="-DVR_SMEM_STACK=16 -DVR_TRI48"
  [tex8]="-DVR_TEX8"
  [stack16_tri48_tex8]="-DVR_SMEM_STACK=16 -DVR_TRI48 -DVR_TEX8"
  [spec]="-DVR_TRACE_SPEC"
  [spec_arrival]="-DVR_TRACE_SPEC -DVR_SPEC_ARRIVAL"
  [spec_arrival_unpark]="-DVR_TRACE_SPEC -DVR_SPEC_ARRIVAL -DVR_SPEC_UNPARK"
  [spec_au_lv1]="-DVR_TRACE_SPEC -DVR_SPEC_ARRIVAL -DVR_SPEC_UNPARK -DVR_LEAF_VOTE_NUM=1"
  [spec_au_lv1_stack16_tex8]="-DVR_TRACE_SPEC -DVR_SPEC_ARRIVAL -DVR_SPEC_UNPARK -DVR_LEAF_VOTE_NUM=1 -DVR_SMEM_STACK=16 -DVR_TEX8"
  [spec_once]="-DVR_TRACE_SPEC -DVR_SPEC_PARK_ONCE"
  [spec_b7]="-DVR_TRACE_SPEC -DVR_TRACE_MIN_BLOCKS=7"
  [spec_lv1_ls4]="-DVR_TRACE_SPEC -DVR_LEAF_VOTE_NUM=1 -DVR_LEAF_STEPS=4"
  [spec_stack16]="-DVR_TRACE_SPEC -DVR_SMEM_STACK=16"
  [spec_stack16_tex8]="-DVR_TRACE_SPEC -DVR_SMEM_STACK=16 -DVR_TEX8"
  [spec_stack16_tri48_tex8]="-DVR_TRACE_SPEC -DVR_SMEM_STACK=16 -DVR_TRI48 -DVR_TEX8"
  [bvh4_spec]="-DVR_BVH4 -DVR_NODE_STEPS=2 -DVR_TRACE_SPEC"
  [chunk]="-DVR_TRACE_CHUNK -DVR_LEAF_COMPACT"
  [chunk_r16]="-DVR_TRACE_CHUNK -DVR_LEAF_COMPACT -DVR_REFILL_THRESHOLD=16"
  [chunk_r20]="-DVR_TRACE_CHUNK -DVR_LEAF_COMPACT -DVR_REFILL_THRESHOLD=20"
  [chunk_r24]="-DVR_TRACE_CHUNK -DVR_LEAF_COMPACT -DVR_REFILL_THRESHOLD=24"
  [chunk_r20_b7]="-DVR_TRACE_CHUNK -DVR_LEAF_COMPACT -DVR_REFILL_THRESHOLD=20 -DVR_TRACE_MIN_BLOCKS=7"
  [bvh4_chunk_r20_b7]="-DVR_BVH4 -DVR_NODE_STEPS=2 -DVR_TRACE_CHUNK -DVR_LEAF_COMPACT -DVR_REFILL_THRESHOLD=20 -DVR_TRACE_MIN_BLOCKS=7"
  [bvh4]="-DVR_BVH4 -DVR_NODE_STEPS=2"
  [bvh4_steps3]="-DVR_BVH4 -DVR_NODE_STEPS=3"
  [bvh4_steps1]="-DVR_BVH4 -DVR_NODE_STEPS=1"
  [bvh4_nosort]="-DVR_BVH4 -DVR_BVH4_NOSORT -DVR_NODE_STEPS=2"
  [bvh4_stack16]="-DVR_BVH4 -DVR_NODE_STEPS=2 -DVR_SMEM_STACK=16"
  [bvh4_stack16_tex8]="-DVR_BVH4 -DVR_NODE_STEPS=2 -DVR_SMEM_STACK=16 -DVR_TEX8"
)
names=("$@")
[ ${#names[@]} -eq 0 ] && names=("${!FLAGS[@]}")
mkdir -p gpurun_variants
for name in "${names[@]}"; do
  flags="${FLAGS[$name]}"
  [ -z "$flags" ] && { echo "unknown variant $name"; exit 2; }
  obj="build/variants/$name"
  mkdir -p "$obj"
  echo "== $name: $flags"
  make -C voidray_b200/csrc -s -j4 OBJDIR="$PWD/$obj" OUT="$PWD/gpurun_variants/$name.so" EXTRA="$flags" 2>&1 |
    grep -A2 "k_traceENS" | grep -E "registers|stack frame" || true
done
ls -la gpurun_variants/
