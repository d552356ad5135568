#!/bin/bash
# Round 2, GPU call AE: the final build end to end: full parity suite, bench.py with configs, scenes table, ncu launch list + full set (Cornell, C4, C5).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ae_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02ae_bench.json 2> gpurun_out/r02ae_bench.err
timeout 600 python tools/bench_scenes.py > gpurun_out/r02ae_scenes.md 2> gpurun_out/r02ae_scenes.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02ae_cornell_launches.csv python tools/profile_step.py cornell 2 > gpurun_out/r02ae_ncu_l.log 2>&1
timeout 900 $NCU --set full --import-source on -c 10 -o /tmp/ncu/cornell_final -f python tools/profile_step.py cornell 1 > gpurun_out/r02ae_ncu_f.log 2>&1
bash tools/ncu_export.sh /tmp/ncu/cornell_final.ncu-rep r02ae_cornell_final 6
for sc in hdri2 instanced_monkeys; do
  timeout 900 $NCU --set full -c 12 -o /tmp/ncu/${sc}_final -f python tools/profile_step.py $sc 1 > gpurun_out/r02ae_ncu_$sc.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/${sc}_final.ncu-rep gpurun_out/r02ae_${sc}_final_kernels.csv > /dev/null 2>&1
done
du -sh gpurun_out
set +x
echo ==== PYTEST; grep -E "relMSE|passed|failed|^FAILED|^E  |rc=|furnace_exact" gpurun_out/r02ae_pytest.log | tail -30
echo ==== SCENES; cat gpurun_out/r02ae_scenes.md; tail -3 gpurun_out/r02ae_scenes.err
echo ==== BENCH; tail -3 gpurun_out/r02ae_bench.err; python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02ae_bench.json").read().strip().splitlines()[-1])
print("value %.3f G, ms %.2f, dev ms %.2f (instr %.2f), e2e %.3f G" % (j["value"]/1e9, j["ms_per_step"], j["device_ms_per_step"], j["device_ms_per_step_instrumented"], j["e2e"]["value"]/1e9))
print(j["e2e"]["rank0_step_ms"], j["roofline"]["kernel"], j["roofline"]["frac"], j["frame_hbm_roofline"])
print(j["kernel_time_share"])
for c in j["configs"]: print("  ", c["id"], c["scene"], c["film"], c["total_spp"], "spp: %.3f Gseg/s, %.1f ms, dom %s %s" % (c["value"]/1e9, c["ms_per_step"], c["dominant_kernel"], c["dominant_kernel_roofline"]))
print(j["cpu_baseline"]); print(j["multi_inprocess"])
PY
