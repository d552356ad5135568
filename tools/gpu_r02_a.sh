#!/bin/bash
# Round 2, GPU call A: parity suite, measured peaks, small-scene mode A/B, ncu evidence for Cornell / C5 / C4.
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 300 python tools/bw_probe.py > gpurun_out/r02_peaks.json 2> gpurun_out/r02_peaks.err
timeout 600 python tools/bench_scenes.py cornell furnace gem hdri instanced_monkeys test_nee_sphere orb_caustic sun_test rtiow2 kitchen_sink > gpurun_out/r02a_scenes.md 2> gpurun_out/r02a_scenes.err
RPT_NO_SMALL=1 timeout 300 python tools/bench_scenes.py cornell furnace hdri test_nee_sphere orb_caustic sun_test rtiow2 > gpurun_out/r02a_scenes_nosmall.md 2>> gpurun_out/r02a_scenes.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
# ncu: launch list of two Cornell passes (time + dram bytes), then full sets
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02a_cornell_launches.csv python tools/profile_step.py cornell 2 > gpurun_out/r02a_ncu_l.log 2>&1
timeout 900 $NCU --set full --import-source on -c 9 -o gpurun_out/r02a_cornell_full -f python tools/profile_step.py cornell 1 > gpurun_out/r02a_ncu_f.log 2>&1
RPT_NO_SMALL=1 timeout 900 $NCU --set full -k regex:"k_trace|k_shadow" -c 4 -o gpurun_out/r02a_cornell_bvh_full -f python tools/profile_step.py cornell 1 > gpurun_out/r02a_ncu_fb.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"k_trace|k_shadow" -c 4 -o gpurun_out/r02a_monkeys_full -f python tools/profile_step.py instanced_monkeys 1 > gpurun_out/r02a_ncu_m.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"k_shade_surface|k_shade_miss" -c 3 -o gpurun_out/r02a_hdri_full -f python tools/profile_step.py hdri 1 > gpurun_out/r02a_ncu_h.log 2>&1
for f in cornell_full cornell_bvh_full monkeys_full hdri_full; do
  python tools/ncu_summary.py gpurun_out/r02a_$f.ncu-rep gpurun_out/r02a_${f}_kernels.csv > /dev/null 2>&1
done
ls -la gpurun_out | tail -30
tail -5 gpurun_out/r02a_pytest.log
cat gpurun_out/r02a_scenes.md gpurun_out/r02a_scenes_nosmall.md
cat gpurun_out/r02_peaks.json
