#!/bin/bash
# Round 2, GPU call L: the shipped build end to end (final): parity suite, bench.py with configs, scene table, ncu evidence.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02l_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err
timeout 600 python tools/bench_scenes.py cornell furnace gem hdri hdri_4k hdri2 instanced_monkeys test_nee_sphere orb_caustic sun_test rtiow2 kitchen_sink > gpurun_out/r02l_scenes.md 2> gpurun_out/r02l_scenes.err
RPT_TMA_TILES=1 timeout 300 python tools/bench_scenes.py cornell instanced_monkeys > gpurun_out/r02l_scenes_tma.md 2>> gpurun_out/r02l_scenes.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02l_cornell_launches.csv python tools/profile_step.py cornell 2 > gpurun_out/r02l_ncu_l.log 2>&1
timeout 900 $NCU --set full --import-source on -c 10 -o /tmp/ncu/cornell_final -f python tools/profile_step.py cornell 1 > gpurun_out/r02l_ncu_f.log 2>&1
bash tools/ncu_export.sh /tmp/ncu/cornell_final.ncu-rep r02l_cornell_final
timeout 900 $NCU --set full -k regex:"k_trace|k_shadow|k_nee|k_shade_vertex" -c 12 -o /tmp/ncu/monkeys_final -f python tools/profile_step.py instanced_monkeys 1 > gpurun_out/r02l_ncu_m.log 2>&1
bash tools/ncu_export.sh /tmp/ncu/monkeys_final.ncu-rep r02l_monkeys_final
timeout 900 $NCU --set full -k regex:"k_nee|k_shade_vertex|k_shade_miss|k_shadow" -c 8 -o /tmp/ncu/hdri2_final -f python tools/profile_step.py hdri2 1 > gpurun_out/r02l_ncu_h.log 2>&1
bash tools/ncu_export.sh /tmp/ncu/hdri2_final.ncu-rep r02l_hdri2_final
du -sh gpurun_out
set +x
cat gpurun_out/r02l_smoke.log | tail -2
echo ==== PYTEST; grep -E "passed|failed|^FAILED|^E  |rc=" gpurun_out/r02l_pytest.log | tail -12
echo ==== SCENES; cat gpurun_out/r02l_scenes.md; echo ==== TMA; cat gpurun_out/r02l_scenes_tma.md; tail -3 gpurun_out/r02l_scenes.err
echo ==== BENCH; tail -3 gpurun_out/r02l_bench.err; cat gpurun_out/r02l_bench_ref.json | cut -c1-400; python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02l_bench.json").read().strip().splitlines()[-1])
print("value %.3f G, ms %.2f, dev ms %.2f (instr %.2f), e2e %.3f G" % (j["value"]/1e9, j["ms_per_step"], j["device_ms_per_step"], j["device_ms_per_step_instrumented"], j["e2e"]["value"]/1e9))
print(j["e2e"]["rank0_step_ms"], j["roofline"]["kernel"], j["roofline"]["frac"], j["frame_hbm_roofline"])
print(j["kernel_time_share"])
for c in j["configs"]: print("  ", c["id"], c["scene"], c["film"], c["total_spp"], "spp: %.3f Gseg/s, %.1f ms, dom %s %s" % (c["value"]/1e9, c["ms_per_step"], c["dominant_kernel"], c["dominant_kernel_roofline"]))
print(j["cpu_baseline"])
PY
