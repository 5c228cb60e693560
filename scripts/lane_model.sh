#!/bin/bash
# Lane occupancy of k_trace without a GPU: builds tests/c/wavefront_host.cpp (the kernels of csrc/kernels.cu on the CPU
# shim) with -DVR_HOST_STATS plus the given variant flags and prints, per wavefront depth, votes, live lanes, lanes per
# executed node / leaf step, rays per refill and the issue-slot model of profiles/README.md.
#   bash scripts/lane_model.sh <name> "<node> <leaf> <vote> <refill>" [flags...]
# The four numbers are the variant's instructions per executed node step, leaf step, vote and refill, read off its
# SASS (distance between consecutive node loads / triangle loads / ballots of k_trace in cuobjdump -sass):
#   shipped 64 104 24 170 | spec 74 112 28 170 | spec_once 65 112 29 170 | bvh4 149 104 23 170 | bvh4_nosort 141 104 23 170
#   | chunk 69 109 24 190
#   SCENE="assets/mushroom.obj" bash scripts/lane_model.sh spec "74 112 28 170" -DVR_TRACE_SPEC
set -e
cd "$(dirname "$0")/.."
name=$1; costs=$2; shift 2
read c_node c_leaf c_vote c_refill <<< "$costs"
out=${TMPDIR:-/tmp}/wf_stats_$name
g++ -O2 -std=c++20 -pthread -ffp-contract=off -DVR_HOST_SHIM -DVR_HOST_SIMT -DVR_HOST_STATS "$@" -Itests/c -Ivoidray_b200/csrc \
    -x c++ voidray_b200/csrc/scene_build.cpp tests/c/wavefront_host.cpp -o "$out"
echo "== $name ($*), ${SCENE:-assets/mossy_ground.obj}, 128 x 96 x 4 spp, the mushroom example's view"
COST_NODE=$c_node COST_LEAF=${c_leaf:-104} COST_VOTE=${c_vote:-24} COST_REFILL=${c_refill:-170} "$out" "${SCENE:-assets/mossy_ground.obj}" 128 96 4 6 0x5EED0001 0.2 2.8 -10.5 0.2 0.8 -0.5 0.17 0.7 0.8 0.7 0.6 0.6 0.6 \
    "${TMPDIR:-/tmp}/wf_out_$name.bin" | sed 's/ lane-steps per ray//'
