#!/bin/bash
# First GPU call of the next session, one gpurun (~15 min of box time for parts 1, 2 and 4, ~3 min more per variant of
# part 3: about 55 min with the 14 variants listed; VARIANTS="stack16 tex8" for a first short pass):
#   bash scripts/build_variants.sh                      # here, before the call (the .so files travel with the snapshot;
#                                                       # they are git-ignored, so a fresh container has to rebuild them)
#   gpurun --timeout 4500 -- 'bash scripts/r2_first_call.sh > gpurun_out/r2_first.log 2>&1'   (in the background)
# 1. the GPU gates on the default build (the commit path changed on the host since the last GPU run: same bytes, pinned
#    by digests on CPU, but this is the first time the device sees them again)
# 2. bench.py on all five configs (refreshes BASELINE.md §5: the host commit is 1.3-2x faster, e2e moves)
# 3. the single-flag experiment variants. The CPU lane model (profiles/README.md) ranks none of the traversal variants
#    below the shipped kernel in issue slots, so the ones that act on latency and cache come first (shared-memory stack,
#    RGBA8 textures, 48-byte triangles, 4-wide nodes), then the parked leaf and the chunked claims:
#    closest-hit + radiance gates, then the three-regime perf check
# 4. launch list of the default build for profiles/
mkdir -p gpurun_out
echo "=== gates"; python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== bench, all configs"; bash scripts/bench_all.sh r2
# (the combined variants stack16_tri48*, bvh4_stack16* wait for the single-flag results)
echo "=== variants"; PARITY=1 VARIANTS="${VARIANTS:-spec_au_lv1 stack16 tex8 tri48 bvh4_nosort bvh4 bvh4_steps1 spec_arrival_unpark spec_arrival spec_once bvh4_spec chunk chunk_r20_b7}" bash scripts/perf_variants.sh
echo "=== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config2_r2.csv \
    python bench.py --workload config2_mossy_ground --spp 16 --steps 1 --warmup 3 --no-cpu --no-extra > gpurun_out/launches_config2_r2.log 2>&1
tail -1 gpurun_out/launches_config2_r2.log | cut -c1-300
