#!/usr/bin/env python3
"""Small renders of every code path for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
p_ = parity.pkg()
for name in sys.argv[1:] or ["cornell", "kitchen_sink", "hdri", "instanced_monkeys"]:
    world, st, flat = parity.load_scene(name, 48, 27, 2)
    cs = parity.cuda_scene(flat)
    f, c = cs.render_pt(st.params(seed=1))
    cs.trace_primary(st.params(seed=1))
    rs = p_.loader.RenderSettings(filename="x", width=48, height=27, integrator_type="PT", light_samples=2, medium_aware=False, min_bounces=1, max_bounces=4,
                                  hwss=False, threads=1, min_samples=2, camera_id="main", russian_roulette=True, only_direct=False, wavelength_bounds=None, premultiply=None)
    rs.raw = {"tonemap_settings": {"type": "Reinhard1", "luminance_only": False, "key_value": 0.18, "white_point": 1.0}, "colorspace_settings": {"type": "sRGB"}}
    cs.output_film(p_.renderer.output_settings(rs), None, 48, 27)
    if world.environment.kind == 2:
        lum, basis = p_.importance_map.bake_curve_tables(world, p_.curves.y_bar_curve(), st.wavelength_bounds)
        cs.bake_importance_map(33, 65, lum, basis, st.wavelength_bounds)
        cs.render_pt(st.params(seed=2))
    print(name, "ok", float(f[..., 1].mean()), c.segments)
    cs.close()
