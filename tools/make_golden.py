#!/usr/bin/env python3
"""Writes the committed golden fixtures under tests/golden/ from the CPU oracle (run here, commit the output)."""
import ctypes as ct, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
G = os.path.join(ROOT, "tests", "golden")
os.makedirs(G, exist_ok=True)
lib = parity.oracle_lib()
F = ct.c_float
lib.rpto_ggx_bsdf.argtypes = [F, F, F, F, ct.c_int, ct.c_void_p, ct.c_void_p, ct.POINTER(F), ct.POINTER(F)]
lib.rpto_philox.argtypes = [ct.c_uint64, ct.c_uint32, ct.c_uint32, ct.c_uint32, ct.c_void_p]
v = lambda a: (F * 3)(*a)
eta = lambda lam: float(np.float32(1.5) + np.float32(10000.0) / (np.float32(lam) * np.float32(lam)))
# fixed direction pairs of the reference's GGX tests (ggx.rs:825-826,888-890,905-906,923-924,933-935)
pairs = [
    (0.001, 500.0, [0.9709351, 0.18724124, 0.14908342], [-0.008856451, 0.6295874, -0.7768792]),
    (0.001, 762.2971, [0.073927574, -0.9872729, 0.1408083], [0.048132252, 0.5836164, -0.81060183]),
    (0.01, 500.0, [0.48507738, 0.4317013, -0.76048267], [-0.7469567, -0.66481555, 0.00871551]),
    (0.01, 500.0, [0.95028764, -0.24520797, 0.19190234], [-0.19736944, 0.961363, -0.19190234]),
    (0.01, 762.2971, [0.073927574, -0.9872729, 0.1408083], [0.048132252, 0.5836164, -0.81060183]),
]
out = []
for alpha, lam, wi, wo in pairs:
    for a, b in ((wi, wo), (wo, wi)):
        f, p = F(), F()
        lib.rpto_ggx_bsdf(alpha, eta(lam), 1.0, 0.0, 0, v(a), v(b), ct.byref(f), ct.byref(p))
        out.append({"alpha": alpha, "lambda": lam, "wi": a, "wo": b, "f": f.value, "pdf": p.value})
json.dump(out, open(os.path.join(G, "ggx_fixed.json"), "w"), indent=1)
ph = []
o4 = (F * 4)()
for seed, pixel, sample, block in [(0, 0, 0, 0), (1, 2, 3, 4), (0xDEADBEEFCAFE, 2073599, 15, 27), (7, 123456, 1023, 2)]:
    lib.rpto_philox(seed, pixel, sample, block, o4)
    ph.append({"seed": seed, "pixel": pixel, "sample": sample, "block": block, "out": list(o4)})
json.dump(ph, open(os.path.join(G, "philox.json"), "w"), indent=1)
world, st, flat = parity.load_scene("cornell", 16, 9, 4)
sc = parity.oracle_scene(flat)
film, _ = sc.render_pt(st.params(seed=11))
np.save(os.path.join(G, "cornell_oracle_16x9.npy"), film)
# one tiny film per scene blob: pins every oracle code path (materials, lights, environments, instancing, cameras)
films = {}
for name in ["cornell", "furnace", "furnace_exact", "gem", "hdri", "hdri2", "test_nee_sphere", "orb_caustic", "sun_test", "rtiow2", "instanced_monkeys", "kitchen_sink"]:
    world, st, flat = parity.load_scene(name, 16, 12, 4)
    sc = parity.oracle_scene(flat)
    films[name], _ = sc.render_pt(st.params(seed=11))
    sc.close()
# the importance-map bake (N3) at a small resolution
world, st, flat = parity.load_scene("hdri", 16, 12, 1)
sc = parity.oracle_scene(flat)
p_ = parity.pkg()
lum, basis = p_.importance_map.bake_curve_tables(world, p_.curves.y_bar_curve(), st.wavelength_bounds)
bk = sc.bake_importance_map(12, 20, lum, basis, st.wavelength_bounds)
films["hdri_imap_row_cdf_12x20"] = bk["row_cdf"]
films["hdri_imap_marginal_cdf_12"] = bk["marginal_cdf"]
sc.close()
np.savez_compressed(os.path.join(G, "oracle_films_16x12.npz"), **films)
print("golden fixtures written to", G)
