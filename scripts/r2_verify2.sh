#!/bin/bash
# Round 2, verification of the committed tree: full gates, then the headline bench line (config 1 only).
mkdir -p gpurun_out
echo "=== gates"; timeout -k 10 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== bench (config 1)"; timeout -k 10 300 python bench.py --no-extra > gpurun_out/bench_r2_verify2.json 2> gpurun_out/bench_r2_verify2.err; tail -c 300 gpurun_out/bench_r2_verify2.err; cut -c1-250 gpurun_out/bench_r2_verify2.json
