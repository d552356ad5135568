#!/bin/bash
# Round 2, GPU call N: where does the per-frame fixed cost come from? Frame time against work size.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r02n_fixed.txt 2> gpurun_out/r02n.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
for (w, h, spp) in ((16, 16, 1), (256, 144, 1), (960, 540, 1), (1920, 1080, 1), (1920, 1080, 2), (1920, 1080, 4), (1920, 1080, 8), (1920, 1080, 16), (1920, 1080, 32), (1920, 1080, 64)):
    world, st, flat = parity.load_scene("cornell", w, h, spp)
    sc = parity.cuda_scene(flat)
    best, bestk = 1e9, None
    for i in range(5):
        ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0))
        if i: best = min(best, c.device_ms)
    ptr, c = sc.render_pt_device(st.params(seed=9, spp_total=0, flags=1))
    kt = {k["name"].replace("k_", ""): round(k["ms"], 3) for k in sc.kernel_times()}
    print(f"{w}x{h} spp {spp:3d}: slots {w*h*spp:10d}  {best:9.3f} ms  {c.segments / best / 1e6:6.3f} Gseg/s  launches {c.kernel_launches}  per-kernel ms (instrumented) {kt}")
    sc.close()
PY
cat gpurun_out/r02n_fixed.txt; tail -3 gpurun_out/r02n.err
