#!/usr/bin/env python3
"""First-contact GPU diagnostic: per-scene hit-id match, image agreement, counters, timings."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity

names = sys.argv[1:] or ["cornell", "furnace", "furnace_exact", "gem", "hdri", "test_nee_sphere", "orb_caustic", "sun_test", "parallel_prism", "lighting_north", "rtiow2", "instanced_monkeys", "kitchen_sink"]
for name in names:
    try:
        w, h = (192, 108)
        world, st, flat = parity.load_scene(name, w, h, 8)
        cs, os_ = parity.cuda_scene(flat), parity.oracle_scene(flat)
        p = st.params(seed=5)
        gi, gp, gt = cs.trace_primary(p); oi, op, ot = os_.trace_primary(p)
        same = (gi == oi) & (gp == op)
        t0 = time.time(); fg, cg = cs.render_pt(p); tg = time.time() - t0
        t0 = time.time(); fo, co = os_.render_pt(p); to = time.time() - t0
        print(f"{name:18s} hit-match {same.mean():.6f} miss g/o {np.mean(gi==0xFFFFFFFF):.3f}/{np.mean(oi==0xFFFFFFFF):.3f} "
              f"Y g/o {fg[...,1].mean():.6f}/{fo[...,1].mean():.6f} relMSE {parity.rel_mse(fg, fo):.3e} finite {np.isfinite(fg).all()} t g/o {tg:.3f}/{to:.3f}s")
        print("    gpu   ", cg.as_dict())
        print("    oracle", co.as_dict())
        if not same.all():
            bad = np.nonzero(~same)[0][:5]
            for b in bad: print("    mismatch pix", b, "gpu", gi[b], gp[b], gt[b], "oracle", oi[b], op[b], ot[b])
        cs.close(); os_.close()
    except Exception as e:
        import traceback; traceback.print_exc()
        print(name, "FAILED", e)
