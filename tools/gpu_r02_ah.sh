#!/bin/bash
# Round 2, GPU call AH (8 GPUs): the final build at 8 GPUs and at 1 GPU on the same box.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02ah_gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02ah_bench_n8.json 2> gpurun_out/r02ah_bench_n8.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02ah_bench_n1.json 2> gpurun_out/r02ah_bench_n1.err
set +x
python - <<'PY'
import json
for n in (1, 8):
    f = f"gpurun_out/r02ah_bench_n{n}.json"
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1200:]); continue
    print(n, "value %.3f G, ms %.2f, dev %.2f, reduce %.3f, e2e %.3f G" % (j["value"]/1e9, j["ms_per_step"], j["device_ms_per_step"], j["reduce_ms_per_step"], j["e2e"]["value"]/1e9), j["e2e"]["rank0_step_ms"])
    for s in j["strong"]: print("   strong", s["total_spp"], "%.3f G %.2f ms reduce %.3f" % (s["value"]/1e9, s["ms_per_step"], s["reduce_ms_per_step"]))
    for c in j["configs"]: print("   ", c["id"], c["scene"], c["total_spp"], "spp: %.3f Gseg/s, %.1f ms, reduce %.2f ms" % (c["value"]/1e9, c["ms_per_step"], c["reduce_ms_per_step"]))
    print("    multi", {k: (round(v["value_e2e"]/1e9, 3), round(v["exchange_device_ms"], 3)) for k, v in (j["multi_inprocess"] or {}).items() if isinstance(v, dict)})
PY
