#!/bin/bash
# compute-sanitizer over one tiny end-to-end pass (smoke(): commit, primary gate, 8 spp wavefront, resolve):
# memcheck (out-of-bounds / misaligned), racecheck (shared-memory stack, warp-aggregated compaction), initcheck.
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/sanitizer_$tool.log | tail -3
done
