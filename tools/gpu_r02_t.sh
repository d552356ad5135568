#!/bin/bash
# Round 2, GPU call T: NEE samples sorted by kind (k_nee<., MIXED>) against the lane-per-vertex loop: equivalence tests,
# same-session A/B on the scenes that mix environment and light samples, ncu rows of k_nee / k_shadow on the GGX + HDR scene.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sorted_by_kind or same_stream or hdri" > gpurun_out/r02t_tests.log 2>&1
tail -5 gpurun_out/r02t_tests.log
timeout 900 python - > gpurun_out/r02t_nee_sort.txt 2> gpurun_out/r02t.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
lib = p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, "librpt_b200.so"))
for name, kw in (("hdri2", {"spp": 32}), ("kitchen_sink", {}), ("instanced_monkeys", {}), ("cornell", {}), ("gem", {"spp": 64})):
    world, st, flat = parity.load_scene(name, **kw)
    for mode in ("0", "1"):
        os.environ["RPT_NEE_SORT"] = mode
        sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
        best = None
        for i in range(4):
            ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
            kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
            if i and (best is None or c.device_ms < best[0]):
                best = (c.device_ms, kt, c)
        ms, kt, c = best
        ks = "  ".join(f"{k.replace('k_', '')} {v:8.2f}" for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:6])
        print(f"{name:18s} NEE_SORT={mode} {ms:9.3f} ms {c.segments / ms / 1e6:6.3f} Gseg/s  {ks}", flush=True)
        sc.close()
PY
cat gpurun_out/r02t_nee_sort.txt; tail -3 gpurun_out/r02t.err
for mode in 0 1; do
  RPT_NEE_SORT=$mode timeout 600 ncu --clock-control none --set full -k regex:"k_nee|k_shadow" -c 3 -o /tmp/ncu/hdri2_sort_$mode -f python tools/profile_step.py hdri2 1 > gpurun_out/r02t_ncu_$mode.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/hdri2_sort_$mode.ncu-rep gpurun_out/r02t_hdri2_nee_sort_${mode}_ncu_kernels.csv > /dev/null 2>&1
done
ls -la gpurun_out/r02t*
