#!/bin/bash
# Round 2, GPU call J: adaptive chunk size (32-entry chunks for launches under 4 M entries) + full parity suite + bench.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
timeout 300 python - > gpurun_out/r02j_chunks.txt 2>> gpurun_out/r02j.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
for name, kw in (("cornell", {}), ("cornell", {"spp": 2}), ("cornell", {"spp": 4}), ("furnace", {}), ("gem", {"spp": 64}), ("instanced_monkeys", {}), ("hdri2", {"spp": 32}), ("kitchen_sink", {}), ("orb_caustic", {}), ("test_nee_sphere", {})):
    world, st, flat = parity.load_scene(name, **kw)
    sc = parity.cuda_scene(flat)
    best = 1e9
    for i in range(5):
        ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0))
        if i: best = min(best, c.device_ms)
    print(f"{name:18s} spp {st.min_samples:4d}: {best:9.3f} ms  {c.segments / best / 1e6:6.3f} Gseg/s  launches {c.kernel_launches}")
    sc.close()
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
set +x
echo ==== PYTEST; tail -6 gpurun_out/r02j_pytest.log
echo ==== CHUNKS; cat gpurun_out/r02j_chunks.txt; tail -3 gpurun_out/r02j.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02j_bench.json").read().strip().splitlines()[-1])
print("value %.3f G, ms %.2f, dev ms %.2f (instr %.2f), e2e %.3f G" % (j["value"]/1e9, j["ms_per_step"], j["device_ms_per_step"], j["device_ms_per_step_instrumented"], j["e2e"]["value"]/1e9), j["e2e"]["rank0_step_ms"])
print(j["strong"]); print(j["kernel_time_share"])
PY
