#!/bin/bash
# Round 2, GPU call M: same-session A/B of the CTA-level counter flush.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r02m_flush.txt 2> gpurun_out/r02m.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
for rep in range(2):
    for so in ("librpt_var_perwarpflush.so", "librpt_b200.so"):
        lib = p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, so))
        for name, kw in (("cornell", {}), ("cornell", {"spp": 2}), ("cornell", {"spp": 128}), ("gem", {"spp": 64}), ("instanced_monkeys", {}), ("kitchen_sink", {}), ("furnace", {})):
            world, st, flat = parity.load_scene(name, **kw)
            sc = p.ffi.Scene(lib, flat, 0)
            best = 1e9
            for i in range(4):
                ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0))
                if i: best = min(best, c.device_ms)
            print(f"{so:26s} {name:18s} spp {st.min_samples:4d}: {best:9.3f} ms  {c.segments / best / 1e6:6.3f} Gseg/s")
            sc.close()
PY
cat gpurun_out/r02m_flush.txt; tail -3 gpurun_out/r02m.err
