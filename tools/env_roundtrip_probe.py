#!/usr/bin/env python3
"""Device comparison of the two uv -> direction -> uv round trips of an unrotated HDR environment (rpt_debug_env_roundtrip):
CUDA libm atan2 / acos vs the libm-free path, each against numpy's libm on the host (what the CPU oracle computes)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
from test_gpu_parity import env_roundtrip_inputs
p = parity.pkg()
lib = p.ffi.load_library()
f32, f64 = np.float32, np.float64
TAU, PI = f32(6.28318530717958647692), f32(3.14159265358979323846)
def host(u, v):
    a, b = ((u - f32(0.5)) * TAU).astype(f32), (v * PI).astype(f32)
    ad, bd = a.astype(f64), b.astype(f64)
    fsp = np.sin(bd).astype(f32)
    x = (fsp * np.cos(ad).astype(f32)).astype(f32)
    y = (fsp * np.sin(ad).astype(f32)).astype(f32)
    zf = np.cos(bd).astype(f32)
    zero = f32(0)  # the identity rotation, applied in f32 like xform_vec does (it only moves the sign of zero components)
    x, y = ((x + zero * y).astype(f32) + zero * zf).astype(f64), ((zero * x + y).astype(f32) + zero * zf).astype(f64)
    z = np.cos(bd).astype(f32).astype(f64)
    th, ph = np.arctan2(y, x).astype(f32), np.arccos(z).astype(f32)
    return (th / f32(2) / PI + f32(0.5)).astype(f32), (ph / PI).astype(f32)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
u, v = env_roundtrip_inputs(np.random.default_rng(3), n)
hu, hv = host(u, v)
lu, lv = p.ffi.debug_env_roundtrip(lib, 0, u, v, fast=False)
fu, fv = p.ffi.debug_env_roundtrip(lib, 0, u, v, fast=True)
print(f"{len(u)} inputs (9 regimes x {n})")
for name, (au, av), (bu, bv) in (("CUDA libm path vs host libm", (lu, lv), (hu, hv)), ("libm-free path vs host libm", (fu, fv), (hu, hv)), ("libm-free path vs CUDA libm path", (fu, fv), (lu, lv))):
    du, dv = au != bu, av != bv
    print(f"{name:34s}: u differs {int(du.sum()):6d}  v differs {int(dv.sum()):6d}")
    for r in range(9):
        s = slice(r * n, (r + 1) * n)
        if du[s].any() or dv[s].any():
            print(f"      regime {r}: u {int(du[s].sum())} v {int(dv[s].sum())}")
