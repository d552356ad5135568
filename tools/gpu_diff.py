#!/usr/bin/env python3
"""Dump GPU and oracle films (same Philox streams) plus per-sample estimates for offline diffing."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity

name = sys.argv[1]
w, h, spp = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
world, st, flat = parity.load_scene(name, w, h, spp)
cs, os_ = parity.cuda_scene(flat), parity.oracle_scene(flat)
p = st.params(seed=5)
fg, _ = cs.render_pt(p); fo, _ = os_.render_pt(p)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"diff_{name}.npz"), gpu=fg, oracle=fo)
d = np.abs(fg[..., 1] - fo[..., 1]); print("max abs diff", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "relMSE", parity.rel_mse(fg, fo))
