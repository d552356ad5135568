"""Scene blobs: a flattened `World` + its render settings in one .npz, so the GPU box (which has
neither /root/reference nor PIL-readable assets) can run the same scenes. Written here by
tools/bake_scenes.py from the reference's own TOML/CSV/OBJ files; read by bench.py, smoke() and
the `-m gpu` tests. Curves are stored as their LUT rows on the bake-time grid.
"""
from __future__ import annotations

import json
from typing import Tuple

import numpy as np

from . import curves as C
from . import world as W

F32 = np.float32


class LutCurve(C.Curve):
    """A curve known only through its samples on a uniform grid (linear interpolation)."""

    def __init__(self, lo: float, hi: float, values: np.ndarray):
        self.lo, self.hi, self.values = float(lo), float(hi), np.asarray(values, dtype=F32)

    def evaluate(self, lam):
        lam = np.asarray(lam, dtype=F32)
        n = len(self.values)
        x = np.clip((lam - F32(self.lo)) / F32(self.hi - self.lo) * F32(n - 1), 0, n - 1)
        snapped = np.round(x)
        x = np.where(np.abs(x - snapped) < 1e-3, snapped, x).astype(F32)  # re-sampling on the stored grid is exact
        i = np.minimum(x.astype(np.int64), n - 2)
        t = (x - i.astype(F32)).astype(F32)
        a, b = self.values[i], self.values[i + 1]
        return (a + t * (b - a)).astype(F32)


def _t3(t):
    return None if t is None else [np.asarray(t.forward).tolist(), np.asarray(t.reverse).tolist()]


def _t3_load(v):
    return None if v is None else W.Transform3(np.asarray(v[0], dtype=np.float64), np.asarray(v[1], dtype=np.float64))


def save_world(path: str, world: W.World, settings: dict, lambda_lo: float, lambda_hi: float, num_lambda: int = 1024) -> None:
    grid = C.lut_grid(lambda_lo, lambda_hi, num_lambda)
    arrays = {"curve_lut": np.stack([c.evaluate_power(grid) for c in world.curves]).astype(F32)}
    meta = {
        "lut": [lambda_lo, lambda_hi, num_lambda],
        "settings": settings,
        "curve_names": world.curve_names,
        "instances": [
            dict(kind=i.kind, origin=list(map(float, i.origin)), size=list(map(float, i.size)), axis=i.axis, two_sided=bool(i.two_sided),
                 mesh=i.mesh, transform=_t3(i.transform), material=int(i.material))
            for i in world.instances
        ],
        "lights": list(map(int, world.lights)),
        "materials": [
            dict(type=m.type, name=m.name, texstack=m.texstack, curve_a=m.curve_a, curve_b=m.curve_b, curve_c=m.curve_c, alpha=m.alpha,
                 sharpness=m.sharpness, sidedness=m.sidedness, metallic=bool(m.metallic))
            for m in world.materials
        ],
        "textures": [dict(channels=t.channels, curves=list(t.curves), recipe=t.recipe) for t in world.textures],
        "texstacks": world.texstacks,
        "env": dict(kind=world.environment.kind, strength=world.environment.strength, curve=world.environment.curve,
                    angular_diameter=world.environment.angular_diameter, sun_direction=list(world.environment.sun_direction),
                    texstack=world.environment.texstack, rotation=_t3(world.environment.rotation),
                    imap_marginal_integral=world.environment.imap_marginal_integral,
                    has_imap=world.environment.imap_row_pdf is not None,
                    # ImportanceMap::Unbaked (importance_map.rs:32-46): resolution only; the tables are baked where the blob is
                    # loaded (on the device by CudaRenderer / tests, naive.rs:469-487). Only the default luminance curve (y_bar).
                    imap_request=None if world.environment.imap_request is None else list(world.environment.imap_request[:2])),
        "env_sampling_probability": world.env_sampling_probability,
        "cameras": [
            dict(name=c.name, origin=c.origin.tolist(), u=c.u.tolist(), v=c.v.tolist(), w=c.w.tolist(), lower_left=c.lower_left.tolist(),
                 horizontal=c.horizontal.tolist(), vertical=c.vertical.tolist(), aperture_diameter=c.aperture_diameter, vfov=c.vfov,
                 focal_distance=c.focal_distance, kind=c.kind, angle_span=[float(c.angle_span[0]), float(c.angle_span[1])])
            for c in world.cameras
        ],
        "camera_names_to_index": world.camera_names_to_index,
        "num_meshes": len(world.meshes),
        "mesh_has_normals": [m.normals is not None for m in world.meshes],
    }
    for i, m in enumerate(world.meshes):
        arrays[f"mesh{i}_v"] = m.vertices
        arrays[f"mesh{i}_i"] = m.indices
        arrays[f"mesh{i}_m"] = m.face_material
        if m.normals is not None:
            arrays[f"mesh{i}_n"] = m.normals
    for i, t in enumerate(world.textures):
        if t.recipe is None:  # (a synthetic map travels as its recipe: a 4096x2048 environment is 134 MB of texels)
            arrays[f"tex{i}"] = t.texels
    e = world.environment
    if e.imap_row_pdf is not None:
        arrays.update(imap_row_pdf=e.imap_row_pdf, imap_row_cdf=e.imap_row_cdf, imap_m_pdf=e.imap_marginal_pdf, imap_m_cdf=e.imap_marginal_cdf)
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)


def load_world(path: str) -> Tuple[W.World, dict, Tuple[float, float, int]]:
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    lo, hi, n = meta["lut"]
    world = W.World()
    lut = z["curve_lut"]
    world.curves = [LutCurve(lo, hi, lut[i]) for i in range(lut.shape[0])]
    world.curve_names = meta["curve_names"]
    for i in range(meta["num_meshes"]):
        world.meshes.append(W.Mesh(z[f"mesh{i}_v"], z[f"mesh{i}_i"], z[f"mesh{i}_n"] if meta["mesh_has_normals"][i] else None, z[f"mesh{i}_m"]))
    for d in meta["instances"]:
        world.instances.append(W.Instance(d["kind"], tuple(d["origin"]), tuple(d["size"]), d["axis"], d["two_sided"], d["mesh"], _t3_load(d["transform"]), d["material"]))
    world.lights = meta["lights"]
    for d in meta["materials"]:
        world.materials.append(W.Material(**d))
    for i, d in enumerate(meta["textures"]):
        recipe = d.get("recipe")
        if recipe is not None:
            from .synth import texels_from_recipe

            world.textures.append(W.Texture(d["channels"], texels_from_recipe(recipe), tuple(d["curves"]), recipe))
        else:
            world.textures.append(W.Texture(d["channels"], z[f"tex{i}"], tuple(d["curves"])))
    world.texstacks = meta["texstacks"]
    e = meta["env"]
    env = W.Environment(kind=e["kind"], strength=e["strength"], curve=e["curve"], angular_diameter=e["angular_diameter"],
                        sun_direction=tuple(e["sun_direction"]), texstack=e["texstack"], rotation=_t3_load(e["rotation"]),
                        imap_marginal_integral=e["imap_marginal_integral"])
    if e["has_imap"]:
        env.imap_row_pdf, env.imap_row_cdf = z["imap_row_pdf"], z["imap_row_cdf"]
        env.imap_marginal_pdf, env.imap_marginal_cdf = z["imap_m_pdf"], z["imap_m_cdf"]
    if e.get("imap_request") is not None and not e["has_imap"]:
        env.imap_request = (int(e["imap_request"][0]), int(e["imap_request"][1]), C.y_bar_curve())
    world.environment = env
    world.env_sampling_probability = meta["env_sampling_probability"]
    for d in meta["cameras"]:
        a = lambda k: np.asarray(d[k], dtype=F32)
        world.cameras.append(W.Camera(d["name"], a("origin"), a("u"), a("v"), a("w"), a("lower_left"), a("horizontal"), a("vertical"),
                                      d["aperture_diameter"], d["vfov"], d["focal_distance"], kind=d.get("kind", 0),
                                      angle_span=tuple(d.get("angle_span", (0.0, 0.0)))))
    world.camera_names_to_index = meta["camera_names_to_index"]
    return world, meta["settings"], (lo, hi, n)
