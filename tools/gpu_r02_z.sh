#!/bin/bash
# Round 2, GPU call Z: compute-sanitizer over the code added since the first sanitizer pass (sorted NEE with its shared-memory
# rings, the light-only NEE kernel, the libm-free environment round trip, the four-wide BVH walk).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py cornell kitchen_sink hdri2 instanced_monkeys sun_test > gpurun_out/r02z_sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/r02z_sanitizer_$tool.log
  RPT_BVH4=1 timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py cornell instanced_monkeys kitchen_sink > gpurun_out/r02z_sanitizer_bvh4_$tool.log 2>&1
  tail -2 gpurun_out/r02z_sanitizer_bvh4_$tool.log
done
