#!/usr/bin/env python3
"""Throughput of every BASELINE config (and the in-tree coverage scenes) at its configured size on one GPU:
samples/s, path segments/s, rays/s (reference definition and true BVH queries), BVH work per ray.
    python tools/bench_scenes.py [scene ...]   -> markdown table on stdout"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity

DEFAULT = ["cornell", "furnace", "gem", "hdri", "instanced_monkeys", "test_nee_sphere", "orb_caustic", "sun_test", "rtiow2", "kitchen_sink"]
SPP_CAP = {"gem": 64}  # 1080p @ 1024 spp is 2.1 G samples: time 64 spp (one wave), the rate is spp-independent
names = sys.argv[1:] or DEFAULT
print("| scene | film | spp | ms | Msamples/s | Gsegments/s | Grays/s (reference def.) | Grays/s (true) | nodes/ray walk | tris/ray walk | insts/ray walk | nodes/ray NEE | top kernels |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for name in names:
    world, st, flat = parity.load_scene(name, spp=SPP_CAP.get(name))
    sc = parity.cuda_scene(flat)
    best = None
    for i in range(3):
        ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=3))
        if best is None or c.device_ms < best[0].device_ms:
            best = (c, sc.kernel_times())
    c, kt = best
    s = c.device_ms / 1e3
    ref_rays = c.camera_rays + c.bounce_rays + c.shadow_rays + c.light_rays
    tot = sum(k["ms"] for k in kt)
    top = ", ".join(f"{k['name'].replace('k_', '')} {100 * k['ms'] / tot:.0f}%" for k in sorted(kt, key=lambda k: -k["ms"])[:3])
    walk = max(1, c.segments)
    nee = max(1, c.shadow_rays_traced)
    print(f"| {name} | {st.width}x{st.height} | {st.min_samples} | {c.device_ms:.1f} | {c.camera_rays / s / 1e6:.0f} | {c.segments / s / 1e9:.2f} | {ref_rays / s / 1e9:.2f} | "
          f"{c.true_rays / s / 1e9:.2f} | {c.walk_nodes / walk:.1f} | {c.walk_tris / walk:.1f} | {c.walk_insts / walk:.1f} | {c.shadow_nodes / nee:.1f} | {top} |")
    sc.close()
