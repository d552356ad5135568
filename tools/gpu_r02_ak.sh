#!/bin/bash
# Round 2, GPU call AK: guide tables for the importance map's CDF inversions (default) vs the plain binary search (RPT_IMAP_GUIDES=0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "imap or hdri or importance or same_stream or converged" > gpurun_out/r02ak_tests.log 2>&1
tail -3 gpurun_out/r02ak_tests.log
timeout 600 python - > gpurun_out/r02ak_imap_guides.txt 2> gpurun_out/r02ak.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
lib = p.ffi.load_library()
for rep in range(2):
    for name, kw in (("hdri2", {"spp": 32}), ("hdri", {"spp": 32})):
        world, st, flat = parity.load_scene(name, **kw)
        for mode in ("0", "1"):
            os.environ["RPT_IMAP_GUIDES"] = mode
            sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
            best = None
            for i in range(4):
                ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
                kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
                if i and (best is None or c.device_ms < best[0]):
                    best = (c.device_ms, kt, c)
            ms, kt, c = best
            ks = "  ".join(f"{k.replace('k_', '')} {v:7.2f}" for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:6])
            print(f"{name:8s} IMAP_GUIDES={mode} {ms:9.3f} ms {c.segments / ms / 1e6:6.3f} Gseg/s  {ks}", flush=True)
            sc.close()
PY
cat gpurun_out/r02ak_imap_guides.txt; tail -3 gpurun_out/r02ak.err
