import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
for name in ("hdri2", "hdri"):
    world, st, flat = parity.load_scene(name, 192, 108, 8)
    films = {}
    for mode in ("0", "1", "1b"):
        os.environ["RPT_ENV_FAST"] = mode[0]
        sc = parity.cuda_scene(flat)
        f, c = sc.render_pt(st.params(seed=37))
        films[mode] = f
        sc.close()
    for a, b in (("0", "1"), ("1", "1b")):
        fa, fb = films[a], films[b]
        d = np.abs(fa - fb)
        rel = d / np.maximum(np.abs(fa), 1e-30)
        bad = ~np.isclose(fa, fb, rtol=1e-5, atol=1e-9)
        print(name, a, "vs", b, "max abs", float(d.max()), "max rel", float(rel.max()), "elements beyond tolerance", int(bad.sum()), "pixels", int(bad.any(axis=2).sum()))
        idx = np.argwhere(bad)[:5]
        for i in idx:
            print("   ", tuple(i), fa[tuple(i)], fb[tuple(i)])
