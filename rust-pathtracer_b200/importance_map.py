"""Host-side bake of the equirectangular importance map (reference src/world/importance_map.rs:78-253).

Startup pre-pass that feeds the device tables of include/rpt.h (`imap_*`); not on the per-ray path.
Row index <-> u, column index <-> v (importance_map.rs:137-140). Each row is a CurveWithCDF with
`pdf`/`cdf` = Curve::Linear{bounds (0,1), Nearest}, normalised per row, pdf_integral = 1
(:158-176); the marginal is `Curve::Linear{row sums / total, Nearest}.to_cdf((0,1), 100)` (:216-244).
"""
from __future__ import annotations

import numpy as np

from . import curves as C
from . import world as W

F32 = np.float32


def _texel_lookup(tex: W.Texture, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Vec2D::at_uv (reference src/vec2d.rs:34-42): nearest texel, uv clamped to [0, 1-eps]."""
    h, w = tex.texels.shape[:2]
    eps = np.finfo(F32).eps
    uu = np.clip(u.astype(F32), F32(0), F32(1) - eps)
    vv = np.clip(v.astype(F32), F32(0), F32(1) - eps)
    x = (uu * F32(w)).astype(np.int64)
    y = (vv * F32(h)).astype(np.int64)
    return tex.texels[y, x]


def bake_importance_map(world: W.World, vertical_resolution: int, horizontal_resolution: int, luminance_curve: C.Curve,
                        wavelength_bounds=C.BOUNDED_VISIBLE_RANGE, num_samples: int = 100) -> None:
    env = world.environment
    stack = world.texstacks[env.texstack]
    R, Cn = vertical_resolution, horizontal_resolution
    lo, hi = wavelength_bounds
    step = (hi - lo) / num_samples
    lam = (lo + step * np.arange(num_samples, dtype=np.float64)).astype(F32)
    lum = luminance_curve.evaluate(lam).astype(np.float64)

    # texel_luminance(uv) = integral over lambda of max(0, lum * max(0, sum_tex sum_c max(0, texel_c * curve_c)))
    # (Curve::Machine clamps each stage at 0; texture.rs:44-77,126-131,246-253). All factors are >= 0 for
    # the shipped basis curves, so the clamps are inert and the integral is linear in the texel channels.
    u = (np.arange(R, dtype=F32) / F32(R))
    v = (np.arange(Cn, dtype=F32) / F32(Cn))
    uu, vv = np.meshgrid(u, v, indexing="ij")
    total = np.zeros((R, Cn), dtype=np.float64)
    for tid in stack:
        tex = world.textures[tid]
        texels = _texel_lookup(tex, uu.ravel(), vv.ravel()).reshape(R, Cn, tex.channels).astype(np.float64)
        for c in range(tex.channels):
            basis = np.maximum(world.curves[tex.curves[c]].evaluate(lam).astype(np.float64), 0.0)
            weight = float(np.sum(lum * basis) * step)
            total += np.maximum(texels[..., c], 0.0) * weight
    row_lum = np.sum(total, axis=1)
    safe = np.where(row_lum == 0.0, 1.0, row_lum)
    env.imap_row_pdf = (total / safe[:, None]).astype(F32)
    env.imap_row_cdf = (np.cumsum(total, axis=1) / safe[:, None]).astype(F32)
    marginal = (row_lum / np.sum(row_lum)).astype(F32)
    cdf = C.Linear(marginal, (0.0, 1.0), "Nearest").to_cdf((0.0, 1.0), 100)
    env.imap_marginal_pdf = marginal
    env.imap_marginal_cdf = cdf.cdf_signal.astype(F32)
    env.imap_marginal_integral = float(cdf.pdf_integral)


def bake_curve_tables(world: W.World, luminance_curve: C.Curve, wavelength_bounds=C.BOUNDED_VISIBLE_RANGE, num_samples: int = 100):
    """The host half of the device bake (include/rpt.h RptImapBake): luminance_curve and every basis curve of the
    environment's texture stack evaluated at lambda_i = lo + i * (hi - lo) / num_samples (Curve::evaluate_integral's
    sample points). -> (luminance (num_samples,), basis (num_textures * 4 * num_samples,))."""
    lo, hi = wavelength_bounds
    step = (hi - lo) / num_samples
    lam = (lo + step * np.arange(num_samples, dtype=np.float64)).astype(F32)
    lum = luminance_curve.evaluate(lam).astype(F32)
    stack = world.texstacks[world.environment.texstack]
    basis = np.zeros((len(stack), 4, num_samples), dtype=F32)
    for k, tid in enumerate(stack):
        tex = world.textures[tid]
        for c in range(tex.channels):
            basis[k, c] = world.curves[tex.curves[c]].evaluate(lam).astype(F32)
    return lum, basis.ravel()


def bake_importance_map_on_device(scene, world: W.World, vertical_resolution: int, horizontal_resolution: int, luminance_curve: C.Curve,
                                  wavelength_bounds=C.BOUNDED_VISIBLE_RANGE, num_samples: int = 100, download: bool = True) -> dict:
    """ImportanceMap::bake_raw through the C ABI (rpt_scene_bake_importance_map): texels never leave the device, the tables
    are installed in `scene`; with download they are also mirrored into world.environment (what the reference caches on disk)."""
    lum, basis = bake_curve_tables(world, luminance_curve, wavelength_bounds, num_samples)
    out = scene.bake_importance_map(vertical_resolution, horizontal_resolution, lum, basis, wavelength_bounds, download)
    if download:
        env = world.environment
        env.imap_row_pdf, env.imap_row_cdf = out["row_pdf"], out["row_cdf"]
        env.imap_marginal_pdf, env.imap_marginal_cdf = out["marginal_pdf"], out["marginal_cdf"]
        env.imap_marginal_integral = out["marginal_integral"]
    return out
