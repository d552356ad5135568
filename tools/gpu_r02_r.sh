#!/bin/bash
# Round 2, GPU call R: four-wide BVH (RPT_BVH4=1) against the two-wide walk: equivalence tests, same-session A/B on the
# configured workload of every bench scene (80-register and 64-register builds of the wide kernels), ncu counter rows on the
# 10 M-triangle instanced scene.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bvh4" > gpurun_out/r02r_tests.log 2>&1
tail -5 gpurun_out/r02r_tests.log
timeout 900 python - > gpurun_out/r02r_bvh4.txt 2> gpurun_out/r02r.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
for name, kw in (("cornell", {}), ("gem", {"spp": 64}), ("instanced_monkeys", {}), ("hdri2", {"spp": 16}), ("kitchen_sink", {}), ("furnace", {}), ("sun_test", {})):
    world, st, flat = parity.load_scene(name, **kw)
    for so, wide in (("librpt_b200.so", "0"), ("librpt_b200.so", "1"), ("librpt_var_wide8.so", "1")):
        os.environ["RPT_BVH4"] = wide
        lib = p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, so))
        sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
        best = None
        for i in range(4):
            ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=3))
            kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
            if i and (best is None or c.device_ms < best[0]):
                best = (c.device_ms, kt, c)
        ms, kt, c = best
        rays = max(c.segments, 1)
        print(f"{name:18s} {so:22s} BVH4={wide} {ms:9.3f} ms {c.segments / ms / 1e6:6.3f} Gseg/s  trace {kt.get('k_trace', 0):8.3f}  shadow {kt.get('k_shadow', 0):8.3f}"
              f"  nodes/ray walk {c.walk_nodes / rays:5.2f} nee {c.shadow_nodes / max(c.shadow_rays_traced, 1):5.2f}", flush=True)
        sc.close()
PY
cat gpurun_out/r02r_bvh4.txt; tail -3 gpurun_out/r02r.err
for wide in 0 1; do
  RPT_BVH4=$wide timeout 600 ncu --clock-control none --set full -k regex:"k_trace|k_shadow" -c 4 -o /tmp/ncu/monkeys_bvh4_$wide -f python tools/profile_step.py instanced_monkeys 1 > gpurun_out/r02r_ncu_$wide.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/monkeys_bvh4_$wide.ncu-rep gpurun_out/r02r_monkeys_bvh4_${wide}_ncu_kernels.csv > /dev/null 2>&1
done
ls -la gpurun_out/r02r*
