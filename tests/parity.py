"""Shared parity helpers: load a scene blob, render it through the CUDA C ABI and through the CPU
oracle (the checker), compare. Used by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
Nothing here reads /root/reference."""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
import sys
from typing import Optional, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

SCENES = os.path.join(ROOT, "scenes")


def pkg():
    return graft.load_package()


def oracle_lib() -> ct.CDLL:
    """The CPU oracle shared library (built on demand with g++; test infrastructure only)."""
    p = pkg()
    path = p.ffi.ORACLE_LIB_PATH
    if not os.path.exists(path):
        graft.build_oracle()
    lib = p.ffi._declare(ct.CDLL(path), "rpto")
    lib.rpto_render_samples.argtypes = [ct.c_void_p, ct.POINTER(p.ffi.RptRenderParams), ct.c_void_p, ct.c_void_p]
    lib.rpto_render_samples.restype = ct.c_int
    return lib


def load_scene(name: str, width: Optional[int] = None, height: Optional[int] = None, spp: Optional[int] = None):
    """-> (world, PTSettings, FlatScene). width/height/spp override the blob's settings; the camera is
    re-aspected exactly like parsing/cameras.rs:191-201 does per render setting."""
    p = pkg()
    world, settings, (lo, hi, n) = p.blob.load_world(os.path.join(SCENES, name + ".npz"))
    st = p.renderer.PTSettings.from_dict(settings)
    if width is not None:
        st.width = width
    if height is not None:
        st.height = height
    if spp is not None:
        st.min_samples = spp
    world.cameras = [c.with_aspect_ratio(st.width / st.height) for c in world.cameras]
    flat = p.ffi.FlatScene(world, st.wavelength_bounds[0], st.wavelength_bounds[1], n)
    return world, st, flat


def _bake_unbaked_importance_map(scene, flat):
    """ImportanceMap::Unbaked -> baked before the first render (reference src/renderer/naive.rs:469-487), through the scene's own
    library: rpt_scene_bake_importance_map on the device, rpto_scene_bake_importance_map in the oracle."""
    env = flat.world.environment
    if env.kind == 2 and env.imap_row_pdf is None and env.imap_request is not None:
        rows, cols, lum = env.imap_request
        pkg().importance_map.bake_importance_map_on_device(scene, flat.world, rows, cols, lum, flat.lambda_bounds, download=False)
    return scene


def cuda_scene(flat, device: int = 0):
    p = pkg()
    return _bake_unbaked_importance_map(p.ffi.Scene(p.ffi.load_library(), flat, device), flat)


def oracle_scene(flat):
    p = pkg()
    return _bake_unbaked_importance_map(p.ffi.Scene(oracle_lib(), flat, 0, "rpto"), flat)


def oracle_samples(scene, params) -> np.ndarray:
    n = params.width * params.height * params.spp
    e = np.zeros(n, dtype=np.float32)
    rc = scene.lib.rpto_render_samples(scene.handle, ct.byref(params), e.ctypes.data_as(ct.c_void_p), None)
    assert rc == 0
    return e.reshape(params.height, params.width, params.spp)


def rel_mse(a: np.ndarray, b: np.ndarray, eps: float = 1e-4) -> float:
    """mean((a-b)^2 / (b^2 + eps)) over XYZ (SURVEY §8c parity metric (c))."""
    a = a[..., :3].astype(np.float64)
    b = b[..., :3].astype(np.float64)
    return float(np.mean((a - b) ** 2 / (b ** 2 + eps)))


def mean_rel_diff(a: np.ndarray, b: np.ndarray) -> float:
    """|mean(a) - mean(b)| / mean(b) on the Y channel."""
    ma, mb = float(a[..., 1].mean()), float(b[..., 1].mean())
    return abs(ma - mb) / max(abs(mb), 1e-12)


def smoke_check(_pkg=None) -> None:
    """One small invocation of the hot path on cuda:0, checked against the oracle."""
    world, st, flat = load_scene("cornell", 160, 90, 4)
    cs, os_ = cuda_scene(flat), oracle_scene(flat)
    params = st.params(seed=7)
    gi, gp, gt = cs.trace_primary(params)
    oi, op, ot = os_.trace_primary(params)
    match = float(np.mean((gi == oi) & (gp == op)))
    film_g, cnt_g = cs.render_pt(params)
    film_o, cnt_o = os_.render_pt(params)
    assert np.isfinite(film_g).all(), "non-finite film"
    assert match >= 0.9999, f"primary hit ids match only {match:.6f}"
    d = mean_rel_diff(film_g, film_o)
    assert d < 2e-3, f"mean Y differs by {d:.3e} from the oracle on identical sample streams"
    assert cnt_g.kernel_launches > 0
    print(f"smoke ok: hit-id match {match:.6f}, mean-Y rel diff {d:.2e}, launches {cnt_g.kernel_launches}, "
          f"segments gpu/oracle {cnt_g.segments}/{cnt_o.segments}")
    cs.close()
    os_.close()
