#!/bin/bash
# Round 2, GPU call W: libm-free uv round trip of the unrotated HDR environment (RPT_ENV_FAST), light-only NEE kernel for
# p_env = 0 scenes: device-path comparison, HDR parity tests, same-session A/B, then the whole GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "env_ or hdri or same_stream or sorted_by_kind or importance" > gpurun_out/r02w_tests.log 2>&1
tail -5 gpurun_out/r02w_tests.log
timeout 900 python - > gpurun_out/r02w_env_fast.txt 2> gpurun_out/r02w.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
lib = p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, "librpt_b200.so"))
for name, kw in (("hdri2", {"spp": 32}), ("hdri", {}), ("cornell", {}), ("gem", {"spp": 64}), ("furnace", {}), ("instanced_monkeys", {}), ("kitchen_sink", {}), ("sun_test", {})):
    world, st, flat = parity.load_scene(name, **kw)
    for mode in (("0", "hdri" in name), ("1", True)):
        if not mode[1]: continue
        os.environ["RPT_ENV_FAST"] = mode[0]
        sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
        best = None
        for i in range(4):
            ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
            kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
            if i and (best is None or c.device_ms < best[0]):
                best = (c.device_ms, kt, c)
        ms, kt, c = best
        ks = "  ".join(f"{k.replace('k_', '')} {v:8.2f}" for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:7])
        print(f"{name:18s} ENV_FAST={mode[0]} {ms:9.3f} ms {c.segments / ms / 1e6:6.3f} Gseg/s  {ks}", flush=True)
        sc.close()
PY
cat gpurun_out/r02w_env_fast.txt; tail -3 gpurun_out/r02w.err
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02w_tests_all.log 2>&1
tail -4 gpurun_out/r02w_tests_all.log
