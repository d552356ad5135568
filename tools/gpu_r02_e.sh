#!/bin/bash
# Round 2, GPU call E: lane-refill traversal A/B + ncu on the instanced scene.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "refill or at_size or primary_hit or random_rays or axis_aligned or same_stream_images" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
RPT_REFILL=0 timeout 600 python tools/bench_scenes.py cornell gem hdri2 instanced_monkeys kitchen_sink sun_test > gpurun_out/r02e_scenes_tile.md 2> gpurun_out/r02e_scenes.err
RPT_REFILL=1 timeout 600 python tools/bench_scenes.py cornell gem hdri2 instanced_monkeys kitchen_sink sun_test > gpurun_out/r02e_scenes_refill.md 2>> gpurun_out/r02e_scenes.err
NCU="ncu --clock-control none"
RPT_REFILL=1 timeout 900 $NCU --set full --import-source on -k regex:"k_trace|k_shadow" -c 4 -o /tmp/ncu/monkeys_refill -f python tools/profile_step.py instanced_monkeys 1 > gpurun_out/r02e_ncu_m.log 2>&1
bash tools/ncu_export.sh /tmp/ncu/monkeys_refill.ncu-rep r02e_monkeys_refill 3
du -sh gpurun_out
set +x
echo ==== PYTEST; grep -E "relMSE|passed|failed|^FAILED|^E  |rc=" gpurun_out/r02e_pytest.log | tail -30
echo ==== TILE; cat gpurun_out/r02e_scenes_tile.md; echo ==== REFILL; cat gpurun_out/r02e_scenes_refill.md; tail -3 gpurun_out/r02e_scenes.err
cat gpurun_out/r02e_monkeys_refill_kernels.csv | cut -c1-60,250-420
