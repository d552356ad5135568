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
import json
import os
import sys

import numpy as np

ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT_)
import __graft_entry__ as _g  # noqa: E402

_g.load_package()
from rust_pathtracer_b200 import exr, synth  # noqa: E402

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


def write_hdr(path: str, rgb: np.ndarray) -> None:
    """Uncompressed (flat) Radiance RGBE."""
    h, w, _ = rgb.shape
    rgbe = synth.rgbe_encode(rgb)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(rgbe.tobytes())


def write_synthetic_map(path: str, width: int, seed: int) -> None:
    """Writes the synthetic environment map as .hdr (RGBE) or .exr (f32) plus the `<file>.recipe.json` sidecar that lets
    the loader / scene blobs regenerate it (rust-pathtracer_b200/synth.py)."""
    rgb = synth.synth_hdr(width, width // 2, seed)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if path.endswith(".exr"):
        exr.write_exr_rgb(path, rgb)
        enc = "f32"
    else:
        write_hdr(path, rgb)
        enc = "rgbe"
    with open(path + ".recipe.json", "w") as f:
        json.dump({"kind": "synth_hdr", "width": width, "height": width // 2, "seed": seed, "encoding": enc}, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hdri", action="store_true", help="also write the synthetic HDR environment maps")
    ap.add_argument("--hdri-width", type=int, default=4096, help="width of the two maps BASELINE config #4 samples (SURVEY Appendix C2: 4096x2048)")
    args = ap.parse_args()
    write_cornell()
    if args.hdri:
        w = args.hdri_width
        write_synthetic_map(os.path.join(FIX, "hdri", "machine_shop_03_4k.hdr"), w, 41)                   # hdri_test.toml (config #4's scene file)
        write_synthetic_map(os.path.join(FIX, "hdri", "kloofendal_43d_clear_puresky_1k.exr"), w, 43)      # hdri_test_2.toml ("low_res_hdri", lib_textures.toml:25-28)
        write_synthetic_map(os.path.join(FIX, "hdri", "kiara_1_dawn_8k.hdr"), 1024, 42)                   # gem scene: p_env = 0, only ever looked up by escaping rays
    print("fixtures written under", FIX)


if __name__ == "__main__":
    main()
