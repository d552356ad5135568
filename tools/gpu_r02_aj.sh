#!/bin/bash
# Round 2, GPU call AJ: shadow records of one vertex reserved next to each other (variant build -DRPT_NEE_PAIRS) vs the shipped
# per-sample append: equality of results on small renders, then same-session timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python - > gpurun_out/r02aj_nee_pairs.txt 2> gpurun_out/r02aj.err <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
libs = {so: p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, so)) for so in ("librpt_b200.so", "librpt_var_pairs.so")}
for name in ("cornell", "gem", "test_nee_sphere"):
    world, st, flat = parity.load_scene(name, 192, 108, 4)
    out = {}
    for so, lib in libs.items():
        sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
        out[so] = sc.render_pt(st.params(seed=43))
        sc.close()
    (f0, c0), (f1, c1) = out.values()
    same = all(getattr(c0, k) == getattr(c1, k) for k in ("segments", "shadow_rays", "shadow_rays_traced", "env_hits", "nee_vertices"))
    print(f"{name}: counters equal {same}, film max rel diff {float(np.max(np.abs(f0 - f1) / np.maximum(np.abs(f0), 1e-20))):.2e}", flush=True)
for rep in range(2):
    for name, kw in (("cornell", {}), ("gem", {"spp": 64}), ("test_nee_sphere", {}), ("orb_caustic", {})):
        world, st, flat = parity.load_scene(name, **kw)
        for so, lib in libs.items():
            sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
            best = None
            for i in range(5):
                ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
                kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
                if i and (best is None or c.device_ms < best[0]):
                    best = (c.device_ms, kt, c)
            ms, kt, c = best
            ks = "  ".join(f"{k.replace('k_', '')} {v:7.2f}" for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:5])
            print(f"{name:16s} {so:20s} {ms:9.3f} ms {c.segments / ms / 1e6:6.3f} Gseg/s  {ks}", flush=True)
            sc.close()
PY
cat gpurun_out/r02aj_nee_pairs.txt; tail -3 gpurun_out/r02aj.err
