#!/bin/bash
# Round 2, GPU call D: full parity suite on the split shade pipeline, split-vs-fused A/B, ncu of the shipped build.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
timeout 600 python tools/bench_scenes.py cornell furnace gem hdri2 instanced_monkeys kitchen_sink orb_caustic > gpurun_out/r02d_scenes_split.md 2> gpurun_out/r02d_scenes.err
RPT_FUSED_SHADE=1 timeout 600 python tools/bench_scenes.py cornell furnace gem hdri2 instanced_monkeys kitchen_sink orb_caustic > gpurun_out/r02d_scenes_fused.md 2>> gpurun_out/r02d_scenes.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02d_cornell_launches.csv python tools/profile_step.py cornell 2 > gpurun_out/r02d_ncu_l.log 2>&1
timeout 900 $NCU --set full --import-source on -c 10 -o /tmp/ncu/cornell_split -f python tools/profile_step.py cornell 1 > gpurun_out/r02d_ncu_f.log 2>&1
bash tools/ncu_export.sh /tmp/ncu/cornell_split.ncu-rep r02d_cornell_split 6 7
du -sh gpurun_out
set +x
echo ==== PYTEST; grep -E "relMSE|passed|failed|^FAILED|^E  |rc=" gpurun_out/r02d_pytest.log | tail -40
echo ==== SPLIT; cat gpurun_out/r02d_scenes_split.md; echo ==== FUSED; cat gpurun_out/r02d_scenes_fused.md; tail -3 gpurun_out/r02d_scenes.err
echo ==== BENCH; tail -3 gpurun_out/r02d_bench.err; python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02d_bench.json").read().strip().splitlines()[-1])
print("value %.3f G, ms %.2f, dev ms %.2f (instr %.2f), e2e %.3f G" % (j["value"]/1e9, j["ms_per_step"], j["device_ms_per_step"], j["device_ms_per_step_instrumented"], j["e2e"]["value"]/1e9))
print(j["e2e"]["rank0_step_ms"], j["roofline"]["kernel"], j["roofline"]["frac"], j["frame_hbm_roofline"])
print(j["kernel_time_share"]); print(j["kernel_rooflines"])
PY
cat gpurun_out/r02d_cornell_split_kernels.csv | cut -c1-330
