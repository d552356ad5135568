#!/bin/bash
# Round 2, GPU call AO: slots of a frame over 8 x 4 pixel tiles (default) vs row-major (RPT_TILED=0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "tiled or same_stream or public_api or at_size or primary" > gpurun_out/r02ao_tests_all.log 2>&1
tail -3 gpurun_out/r02ao_tests_all.log
timeout 900 python - > gpurun_out/r02ao_tiled_slots.txt 2> gpurun_out/r02ao.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
lib = p.ffi.load_library()
for rep in range(2):
    for name, kw in (("cornell", {}), ("instanced_monkeys", {}), ("hdri2", {"spp": 16}), ("furnace", {})):
        world, st, flat = parity.load_scene(name, **kw)
        sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
        for mode in ("0", "1"):
            os.environ["RPT_TILED"] = mode
            best = None
            for i in range(4):
                ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
                kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
                if i and (best is None or c.device_ms < best[0]):
                    best = (c.device_ms, kt, c)
            ms, kt, c = best
            ks = "  ".join(f"{k.replace('k_', '')} {v:7.2f}" for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:5])
            print(f"{name:18s} TILED={mode} {ms:9.3f} ms {c.segments / ms / 1e6:6.3f} Gseg/s  {ks}", flush=True)
        sc.close()
PY
cat gpurun_out/r02ao_tiled_slots.txt; tail -3 gpurun_out/r02ao.err
