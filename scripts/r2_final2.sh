#!/bin/bash
# Round 2, last call: the final tree (camera-ray culling switched by the host) — instruction counts of the shipped kernels
# on every config for the roofline, the bench line as the driver runs it, smoke.
mkdir -p gpurun_out
echo "=== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "=== instruction counts"; timeout -k 10 900 bash scripts/ncu_trace_inst.sh
echo "=== bench"; timeout -k 10 600 python bench.py > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err; tail -c 400 gpurun_out/bench_r2_final2.err; cut -c1-300 gpurun_out/bench_r2_final2.json
