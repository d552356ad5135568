"""Procedural stand-ins for the environment maps the reference does not ship (the whole
`data/hdri/` directory is missing upstream). A synthetic map is fully described by a small *recipe*
(`{"kind": "synth_hdr", "width", "height", "seed", "encoding"}`), so a scene blob can carry a 4096x2048 environment
(134 MB of texels, the L2-straddling gather BASELINE config #4 is about) as a few bytes and regenerate it wherever it is
loaded. tools/make_fixtures.py writes the same images as .hdr / .exr files under the filenames
reference data/lib_textures.toml names, next to a `<file>.recipe.json` sidecar the loader picks up."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

F32 = np.float32
_CACHE: Dict[Tuple, np.ndarray] = {}


def synth_hdr(width: int, height: int, seed: int) -> np.ndarray:
    """Vertical sky gradient + a 1.5 degree sun disc at 5e4 + three soft windows + low-amplitude value noise. (H, W, 3) float32."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height, dtype=F32) + 0.5) / height
    u = (np.arange(width, dtype=F32) + 0.5) / width
    sky = (2.0 - 1.8 * v)[:, None] * np.ones((1, width), dtype=F32)
    img = np.stack([sky * 0.8, sky * 0.9, sky * 1.1], axis=2).astype(F32)
    theta = (u[None, :] - 0.5) * 2 * np.pi
    phi = v[:, None] * np.pi
    d = np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi) * np.ones_like(theta)], axis=2)
    st, sp = (0.7 - 0.5) * 2 * np.pi, 0.25 * np.pi
    sd = np.array([np.sin(sp) * np.cos(st), np.sin(sp) * np.sin(st), np.cos(sp)], dtype=F32)
    cosang = (d @ sd).astype(F32)
    img[cosang > np.cos(np.deg2rad(1.5))] = np.array([5e4, 4.6e4, 4e4], dtype=F32)
    for (u0, u1, v0, v1, val) in ((0.05, 0.15, 0.40, 0.55, 200.0), (0.30, 0.36, 0.45, 0.60, 50.0), (0.85, 0.95, 0.35, 0.50, 120.0)):
        mu = np.clip(np.minimum(u - u0, u1 - u) / 0.01, 0, 1)
        mv = np.clip(np.minimum(v - v0, v1 - v) / 0.01, 0, 1)
        img += (mv[:, None] * mu[None, :])[..., None] * val
    coarse = rng.random((height // 32 + 1, width // 32 + 1)).astype(F32)
    noise = np.kron(coarse, np.ones((32, 32), dtype=F32))[:height, :width]
    img *= (0.9 + 0.2 * noise)[..., None]
    return img.astype(F32)


def rgbe_encode(rgb: np.ndarray) -> np.ndarray:
    """float RGB -> Radiance RGBE bytes (H, W, 4), the quantisation tools/make_fixtures.py writes into .hdr files."""
    m = np.max(rgb, axis=2)
    _, exp = np.frexp(m)
    scale = np.where(m > 1e-32, 256.0 / np.ldexp(1.0, exp), 0.0).astype(np.float64)
    rgbe = np.zeros(rgb.shape[:2] + (4,), dtype=np.uint8)
    rgbe[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, exp + 128, 0).astype(np.uint8)
    return rgbe


def rgbe_decode(rgbe: np.ndarray) -> np.ndarray:
    """RGBE bytes -> float32 RGB exactly as loader._read_hdr decodes them."""
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e == 0, 0.0, np.ldexp(1.0, e - 136)).astype(F32)
    return (rgbe[..., :3].astype(F32) * scale[..., None]).astype(F32)


def texels_from_recipe(recipe: dict) -> np.ndarray:
    """-> (H, W, 4) float32 RGBA texels of a Texture4 (alpha = recipe['alpha_fill']), cached per process."""
    if recipe.get("kind") != "synth_hdr":
        raise ValueError(f"unknown texture recipe {recipe!r}")
    key = (recipe["width"], recipe["height"], recipe["seed"], recipe.get("encoding", "f32"), float(recipe.get("alpha_fill", 0.0)))
    if key not in _CACHE:
        rgb = synth_hdr(recipe["width"], recipe["height"], recipe["seed"])
        if recipe.get("encoding") == "rgbe":
            rgb = rgbe_decode(rgbe_encode(rgb))
        alpha = np.full(rgb.shape[:2] + (1,), F32(recipe.get("alpha_fill", 0.0)), dtype=F32)
        _CACHE[key] = np.ascontiguousarray(np.concatenate([rgb, alpha], axis=2), dtype=F32)
    return _CACHE[key]
