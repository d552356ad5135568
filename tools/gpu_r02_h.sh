#!/bin/bash
# Round 2, GPU call H (8 GPUs): multi-device tests through the C ABI, bench.py under torchrun at N = 8 and N = 4.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpu.txt
nvidia-smi topo -m > gpurun_out/r02h_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "multi_device" > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02h_bench_n8.json 2> gpurun_out/r02h_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02h_bench_n4.json 2> gpurun_out/r02h_bench_n4.err
set +x
echo ==== PYTEST; tail -4 gpurun_out/r02h_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/r02h_bench_n8.json", "gpurun_out/r02h_bench_n4.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:]); continue
    print(f, "value %.3f G, ms %.2f, dev ms %.2f, reduce %.3f, e2e %.3f G" % (j["value"]/1e9, j["ms_per_step"], j["device_ms_per_step"], j["reduce_ms_per_step"], j["e2e"]["value"]/1e9))
    print("  e2e step ms", j["e2e"]["rank0_step_ms"])
    for s in j["strong"]: print("  strong", s)
    for c in j["configs"]: print("  ", c["id"], c["scene"], c["film"], c["total_spp"], "spp: %.3f Gseg/s, %.1f ms, reduce %.2f ms" % (c["value"]/1e9, c["ms_per_step"], c["reduce_ms_per_step"]))
    print("  multi", j["multi_inprocess"])
PY
