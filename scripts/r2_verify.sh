#!/bin/bash
# Round 2, last call: the final tree — full gates, smoke, the bench line as the driver runs it, the reference arm.
mkdir -p gpurun_out
echo "=== gates"; timeout -k 10 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "=== bench"; timeout -k 10 900 python bench.py > gpurun_out/bench_r2_verify.json 2> gpurun_out/bench_r2_verify.err; tail -c 400 gpurun_out/bench_r2_verify.err; cut -c1-300 gpurun_out/bench_r2_verify.json
echo "=== reference arm"; timeout -k 10 600 python bench.py --impl reference 2>&1 | tail -1 | cut -c1-300
