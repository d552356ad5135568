#!/usr/bin/env python3
"""Synthesises the fixtures the reference tree does not ship:

* fixtures/data/meshes/cornell_box.obj/.mtl — the Cornell box geometry referenced by
  reference data/lib_meshes.toml:1-2 (`*.obj` is git-ignored upstream). Rebuilt from the public
  Cornell measurements (mm) with the axis map the reference scene proves:
  ref = (cornell_z, cornell_x, cornell_y) * 0.001 (camera look_from (-0.8, 0.278, 0.273),
  reference data/scenes/cornell_box.toml:34-35).
* fixtures/data/hdri/*.hdr — seed-fixed synthetic Radiance HDR environment maps under the exact
  filenames reference data/lib_textures.toml names (written only with --hdri; they are large and
  are regenerated on demand rather than committed).

Deterministic; safe to re-run.
"""
import argparse
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "fixtures", "data")

WHITE = [
    # floor, ceiling, back wall
    [(552.8, 0, 0), (0, 0, 0), (0, 0, 559.2), (549.6, 0, 559.2)],
    [(556, 548.8, 0), (556, 548.8, 559.2), (0, 548.8, 559.2), (0, 548.8, 0)],
    [(549.6, 0, 559.2), (0, 0, 559.2), (0, 548.8, 559.2), (556, 548.8, 559.2)],
    # short block
    [(130, 165, 65), (82, 165, 225), (240, 165, 272), (290, 165, 114)],
    [(290, 0, 114), (290, 165, 114), (240, 165, 272), (240, 0, 272)],
    [(130, 0, 65), (130, 165, 65), (290, 165, 114), (290, 0, 114)],
    [(82, 0, 225), (82, 165, 225), (130, 165, 65), (130, 0, 65)],
    [(240, 0, 272), (240, 165, 272), (82, 165, 225), (82, 0, 225)],
    # tall block
    [(423, 330, 247), (265, 330, 296), (314, 330, 456), (472, 330, 406)],
    [(423, 0, 247), (423, 330, 247), (472, 330, 406), (472, 0, 406)],
    [(472, 0, 406), (472, 330, 406), (314, 330, 456), (314, 0, 456)],
    [(314, 0, 456), (314, 330, 456), (265, 330, 296), (265, 0, 296)],
    [(265, 0, 296), (265, 330, 296), (423, 330, 247), (423, 0, 247)],
]
GREEN = [[(0, 0, 559.2), (0, 0, 0), (0, 548.8, 0), (0, 548.8, 559.2)]]
RED = [[(552.8, 0, 0), (549.6, 0, 559.2), (556, 548.8, 559.2), (556, 548.8, 0)]]


def write_cornell():
    os.makedirs(os.path.join(FIX, "meshes"), exist_ok=True)
    lines = ["# synthesised Cornell box (tools/make_fixtures.py); units: metres, ref = (cz, cx, cy) * 0.001",
             "mtllib cornell_box.mtl"]
    nv = 0
    for name, mat, quads in (("cornell_white", "lambertian_white", WHITE), ("cornell_red", "lambertian_red", RED),
                             ("cornell_green", "lambertian_green", GREEN)):
        lines.append(f"o {name}")
        lines.append(f"usemtl {mat}")
        faces = []
        for q in quads:
            for (cx, cy, cz) in q:
                lines.append("v %.6f %.6f %.6f" % (cz * 0.001, cx * 0.001, cy * 0.001))
            faces.append((nv + 1, nv + 2, nv + 3, nv + 4))
            nv += 4
        for f in faces:
            lines.append("f %d %d %d %d" % f)
    with open(os.path.join(FIX, "meshes", "cornell_box.obj"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(FIX, "meshes", "cornell_box.mtl"), "w") as f:
        for mat in ("lambertian_white", "lambertian_red", "lambertian_green"):
            f.write(f"newmtl {mat}\nKd 0.8 0.8 0.8\n\n")


def synth_hdr(width: int, height: int, seed: int) -> np.ndarray:
    """Vertical sky gradient + sun disc + three soft windows + low-amplitude value noise (float32 RGB)."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height, dtype=np.float32) + 0.5) / height
    u = (np.arange(width, dtype=np.float32) + 0.5) / width
    sky = (2.0 - 1.8 * v)[:, None] * np.ones((1, width), dtype=np.float32)
    img = np.stack([sky * 0.8, sky * 0.9, sky * 1.1], axis=2).astype(np.float32)
    # sun: radius 1.5 degrees at (u, v) = (0.7, 0.25)
    theta = (u[None, :] - 0.5) * 2 * np.pi
    phi = v[:, None] * np.pi
    d = np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi) * np.ones_like(theta)], axis=2)
    st, sp = (0.7 - 0.5) * 2 * np.pi, 0.25 * np.pi
    sd = np.array([np.sin(sp) * np.cos(st), np.sin(sp) * np.sin(st), np.cos(sp)], dtype=np.float32)
    cosang = (d @ sd).astype(np.float32)
    img[cosang > np.cos(np.deg2rad(1.5))] = np.array([5e4, 4.6e4, 4e4], dtype=np.float32)
    for (u0, u1, v0, v1, val) in ((0.05, 0.15, 0.40, 0.55, 200.0), (0.30, 0.36, 0.45, 0.60, 50.0), (0.85, 0.95, 0.35, 0.50, 120.0)):
        mu = np.clip(np.minimum(u - u0, u1 - u) / 0.01, 0, 1)
        mv = np.clip(np.minimum(v - v0, v1 - v) / 0.01, 0, 1)
        img += (mv[:, None] * mu[None, :])[..., None] * val
    coarse = rng.random((height // 32 + 1, width // 32 + 1)).astype(np.float32)
    noise = np.kron(coarse, np.ones((32, 32), dtype=np.float32))[:height, :width]
    img *= (0.9 + 0.2 * noise)[..., None]
    return img.astype(np.float32)


def write_hdr(path: str, rgb: np.ndarray) -> None:
    """Uncompressed (flat) Radiance RGBE."""
    h, w, _ = rgb.shape
    m = np.max(rgb, axis=2)
    mant, exp = np.frexp(m)
    scale = np.where(m > 1e-32, 256.0 / np.ldexp(1.0, exp), 0.0).astype(np.float64)
    rgbe = np.zeros((h, w, 4), dtype=np.uint8)
    rgbe[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, exp + 128, 0).astype(np.uint8)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(rgbe.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hdri", action="store_true", help="also write the synthetic HDR environment maps")
    ap.add_argument("--hdri-width", type=int, default=4096)
    args = ap.parse_args()
    write_cornell()
    if args.hdri:
        w = args.hdri_width
        write_hdr(os.path.join(FIX, "hdri", "machine_shop_03_4k.hdr"), synth_hdr(w, w // 2, 41))
        write_hdr(os.path.join(FIX, "hdri", "kiara_1_dawn_8k.hdr"), synth_hdr(w, w // 2, 42))
    print("fixtures written under", FIX)


if __name__ == "__main__":
    main()
