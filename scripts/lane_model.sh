#!/bin/bash
# Lane occupancy of k_trace without a GPU: builds tests/c/wavefront_host.cpp (the kernels of csrc/kernels.cu on the CPU
# shim) with -DVR_HOST_STATS plus the given variant flags and prints, per wavefront depth, votes, live lanes, lanes per
# executed node / leaf step, rays per refill and the issue-slot model of profiles/README.md.
#   bash scripts/lane_model.sh <name> <instructions per node step: 58 BVH2, 140 4-wide, 132 4-wide nosort> [flags...]
#   SCENE="assets/mushroom.obj" bash scripts/lane_model.sh spec 58 -DVR_TRACE_SPEC
set -e
cd "$(dirname "$0")/.."
name=$1; cost=$2; shift 2
out=${TMPDIR:-/tmp}/wf_stats_$name
g++ -O2 -std=c++20 -pthread -ffp-contract=off -DVR_HOST_SHIM -DVR_HOST_SIMT -DVR_HOST_STATS "$@" -Itests/c -Ivoidray_b200/csrc \
    -x c++ voidray_b200/csrc/scene_build.cpp tests/c/wavefront_host.cpp -o "$out"
echo "== $name ($*), ${SCENE:-assets/mossy_ground.obj}, 128 x 96 x 4 spp, the mushroom example's view"
COST_NODE=$cost "$out" "${SCENE:-assets/mossy_ground.obj}" 128 96 4 6 0x5EED0001 0.2 2.8 -10.5 0.2 0.8 -0.5 0.17 0.7 0.8 0.7 0.6 0.6 0.6 \
    "${TMPDIR:-/tmp}/wf_out_$name.bin" | sed 's/ lane-steps per ray//'
