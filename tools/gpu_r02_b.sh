#!/bin/bash
# Round 2, GPU call B: parity suite, measured peaks, small-scene mode A/B, ncu evidence for Cornell / C5 / C4.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
nvidia-smi -L > gpurun_out/r02b_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 300 python tools/bw_probe.py > gpurun_out/r02_peaks.json 2> gpurun_out/r02_peaks.err
timeout 900 python tools/bench_scenes.py cornell furnace gem hdri hdri_4k hdri2 instanced_monkeys test_nee_sphere orb_caustic sun_test rtiow2 kitchen_sink > gpurun_out/r02b_scenes.md 2> gpurun_out/r02b_scenes.err
timeout 300 python tools/bench_scenes.py cornell furnace hdri2 test_nee_sphere orb_caustic sun_test rtiow2 > gpurun_out/r02b_scenes_nosmall.md 2>> gpurun_out/r02b_scenes.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02b_cornell_launches.csv python tools/profile_step.py cornell 2 > gpurun_out/r02b_ncu_l.log 2>&1
timeout 900 $NCU --set full --import-source on -c 9 -o /tmp/ncu/cornell_full -f python tools/profile_step.py cornell 1 > gpurun_out/r02b_ncu_f.log 2>&1
timeout 900 $NCU --set full -k regex:"k_trace|k_shadow" -c 4 -o /tmp/ncu/cornell_bvh_full -f python tools/profile_step.py cornell 1 > gpurun_out/r02b_ncu_fb.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"k_trace|k_shadow" -c 4 -o /tmp/ncu/monkeys_full -f python tools/profile_step.py instanced_monkeys 1 > gpurun_out/r02b_ncu_m.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"k_shade_surface|k_shade_miss" -c 3 -o /tmp/ncu/hdri2_full -f python tools/profile_step.py hdri2 1 > gpurun_out/r02b_ncu_h.log 2>&1
bash tools/ncu_export.sh /tmp/ncu/cornell_full.ncu-rep r02b_cornell_full 1 4
bash tools/ncu_export.sh /tmp/ncu/cornell_bvh_full.ncu-rep r02b_cornell_bvh_full
bash tools/ncu_export.sh /tmp/ncu/monkeys_full.ncu-rep r02b_monkeys_full 2 3
bash tools/ncu_export.sh /tmp/ncu/hdri2_full.ncu-rep r02b_hdri2_full 0
du -sh gpurun_out
set +x
echo ==== PYTEST; tail -15 gpurun_out/r02b_pytest.log
echo ==== SCENES; cat gpurun_out/r02b_scenes.md; echo ==== NOSMALL; cat gpurun_out/r02b_scenes_nosmall.md; tail -3 gpurun_out/r02b_scenes.err
echo ==== BENCH; cut -c1-1500 gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
