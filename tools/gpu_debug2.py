#!/usr/bin/env python3
"""Find the samples where GPU and oracle disagree (spp = 1 => pixel == slot) and print both sides' traces."""
import sys, os, ctypes as ct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
p_ = parity.pkg()
name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
nshow = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w, h = 384, 216
world, st, flat = parity.load_scene(name, w, h, 1)
dbg_lib = p_.ffi.load_library(os.path.join(p_.ffi.PKG_DIR, "librpt_b200_debug.so"))
cs = p_.ffi.Scene(dbg_lib, flat, 0)
os_ = parity.oracle_scene(flat)
p = st.params(seed=5)
fg, cg = cs.render_pt(p); fo, co = os_.render_pt(p)
print("segments", cg.segments, co.segments, "shadow", cg.shadow_rays, co.shadow_rays)
yg, yo = fg[..., 1].ravel(), fo[..., 1].ravel()
diff = np.abs(yg - yo) > 1e-4 * np.maximum(np.abs(yo), 1e-6)
print("differing pixels:", diff.sum(), "of", diff.size, "rows histogram:", np.bincount(np.nonzero(diff)[0] // w, minlength=h)[:20], "...")
lib = os_.lib
lib.rpto_debug_color.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_uint32, ct.c_uint32, ct.c_uint32]
lib.rpto_debug_color.restype = ct.c_float
bad = np.nonzero(diff)[0]
rng = np.random.default_rng(0)
for pix in rng.choice(bad, size=min(nshow, len(bad)), replace=False):
    px, py = int(pix % w), int(pix // w)
    print(f"===== pixel ({px},{py}) slot {pix}: Y gpu {yg[pix]:.7g} oracle {yo[pix]:.7g}")
    sys.stdout.flush()
    lib.rpto_debug_color(os_.handle, ct.byref(p), px, py, 0)
    sys.stdout.flush()
    dbg_lib.rpt_debug_set_slot(ct.c_uint32(int(pix)))
    cs.render_pt(p)
    sys.stdout.flush()
