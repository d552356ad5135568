#!/bin/bash
# Round 2, GPU call C (2 GPUs): the in-library multi-GPU path, the at-size / converged parity tests, bench.py under torchrun.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02c_gpu.txt
nvidia-smi topo -m > gpurun_out/r02c_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_device or at_size or converged or small_scene or reference_parameter or kernel_times" -s > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
set +x
echo ==== PYTEST; grep -E "relMSE|passed|failed|Error|rc=" gpurun_out/r02c_pytest.log | tail -30
echo ==== BENCH N2; tail -5 gpurun_out/r02c_bench_n2.err; python - <<'PY'
import json
for f in ("gpurun_out/r02c_bench_n1.json", "gpurun_out/r02c_bench_n2.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.3f G, ms %.2f, dev ms %.2f (instr %.2f), reduce %.3f, e2e %.3f G" % (j["value"]/1e9, j["ms_per_step"], j["device_ms_per_step"], j["device_ms_per_step_instrumented"], j["reduce_ms_per_step"], j["e2e"]["value"]/1e9))
    print("  e2e step ms", j["e2e"]["rank0_step_ms"], "roofline", j["roofline"]["kernel"], round(j["roofline"]["frac"],3), "frame", round(j["frame_hbm_roofline"]["frac"],3))
    for s in j["strong"]: print("  strong", s)
    for c in j["configs"]: print("  ", c["id"], c["scene"], c["film"], c["total_spp"], "spp: %.3f Gseg/s, %.1f ms, reduce %.2f ms, dom %s %s" % (c["value"]/1e9, c["ms_per_step"], c["reduce_ms_per_step"], c["dominant_kernel"], c["dominant_kernel_roofline"]))
    print("  multi", j["multi_inprocess"])
    print("  cpu", j["cpu_baseline"])
PY
