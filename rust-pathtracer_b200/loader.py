"""Config + scene loading: consumes the reference's TOML / CSV / OBJ / PNG / HDR formats unchanged.

Mirrors reference src/parsing/{config.rs,mod.rs,curves.rs,material.rs,texture.rs,environment.rs,
meshes.rs,instance.rs,primitives.rs,cameras.rs}. It exists only because the Rust host cannot be
built in this image; a Rust `CudaRenderer` flattens a live `World` instead.
Startup code: nothing here is on the per-ray path.

Paths inside the TOML files are relative to the reference crate root ("data/..."); they are
resolved against `roots` in order, so synthesised fixtures (fixtures/data/...) overlay the
reference tree.
"""
from __future__ import annotations

import os
import tomllib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import curves as C
from . import world as W

F32 = np.float32
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_ROOTS = [os.path.join(REPO_ROOT, "fixtures"), "/root/reference"]


class LoadError(RuntimeError):
    pass


class Resolver:
    def __init__(self, roots: Optional[List[str]] = None):
        self.roots = list(roots) if roots else list(DEFAULT_ROOTS)

    def path(self, rel: str) -> str:
        if os.path.isabs(rel) and os.path.exists(rel):
            return rel
        for r in self.roots:
            p = os.path.join(r, rel)
            if os.path.exists(p):
                return p
        raise LoadError(f"could not find {rel!r} under any of {self.roots}")

    def text(self, rel: str) -> str:
        with open(self.path(rel), "r") as f:
            return f.read()

    def toml(self, rel: str) -> dict:
        with open(self.path(rel), "rb") as f:
            return tomllib.load(f)


# ---- config (reference src/parsing/config.rs) ------------------------------------------------


@dataclass
class RenderSettings:
    """reference src/parsing/config.rs:45-62 (+ IntegratorKind :17-31)."""

    filename: Optional[str]
    width: int
    height: int
    integrator_type: str
    light_samples: int
    medium_aware: bool
    min_bounces: Optional[int]
    max_bounces: Optional[int]
    hwss: bool
    threads: Optional[int]
    min_samples: int
    camera_id: str
    russian_roulette: Optional[bool]
    only_direct: Optional[bool]
    wavelength_bounds: Optional[Tuple[float, float]]
    premultiply: Optional[float]
    raw: dict = field(default_factory=dict)


@dataclass
class Config:
    """reference src/parsing/config.rs:132-164."""

    scene_file: str
    renderer: dict
    render_settings: List[RenderSettings]
    env_sampling_probability: Optional[float] = None
    camera_names_to_index: Dict[str, int] = field(default_factory=dict)


_RS_KEYS = {
    "filename", "resolution", "integrator", "min_bounces", "max_bounces", "hwss", "threads", "min_samples",
    "exposure", "max_samples", "camera_id", "russian_roulette", "only_direct", "wavelength_bounds", "premultiply",
    "colorspace_settings", "tonemap_settings",
}


def parse_config(data: dict) -> Config:
    for k in data:
        if k not in ("env_sampling_probability", "default_scene_file", "renderer", "render_settings"):
            raise LoadError(f"unknown config field {k!r} (deny_unknown_fields, config.rs:123)")
    settings = []
    for rs in data["render_settings"]:
        for k in rs:
            if k not in _RS_KEYS:
                raise LoadError(f"unknown render_settings field {k!r} (deny_unknown_fields, config.rs:64)")
        integ = rs["integrator"]
        wb = rs.get("wavelength_bounds")
        settings.append(
            RenderSettings(
                filename=rs.get("filename"),
                width=int(rs["resolution"]["width"]),
                height=int(rs["resolution"]["height"]),
                integrator_type=integ["type"],
                light_samples=int(integ.get("light_samples", 0)),
                medium_aware=bool(integ.get("medium_aware", False)),
                min_bounces=rs.get("min_bounces"),
                max_bounces=rs.get("max_bounces"),
                hwss=bool(rs["hwss"]),
                threads=rs.get("threads"),
                min_samples=int(rs["min_samples"]),
                camera_id=rs["camera_id"],
                russian_roulette=rs.get("russian_roulette"),
                only_direct=rs.get("only_direct"),
                wavelength_bounds=(float(wb[0]), float(wb[1])) if wb else None,
                premultiply=rs.get("premultiply"),
                raw=rs,
            )
        )
    return Config(
        scene_file=data["default_scene_file"],
        renderer=data["renderer"],
        render_settings=settings,
        env_sampling_probability=data.get("env_sampling_probability"),
    )


def get_config(path: str, resolver: Optional[Resolver] = None) -> Config:
    """reference src/parsing/mod.rs:565-581."""
    resolver = resolver or Resolver()
    cfg = parse_config(resolver.toml(path))
    ncpu = os.cpu_count() or 1
    for rs in cfg.render_settings:
        if rs.threads is None:
            rs.threads = ncpu
    return cfg


# ---- OBJ (tobj, single_index + triangulate; reference src/parsing/meshes.rs:17-157) ---------


def load_obj_models(resolver: Resolver, filename: str) -> Tuple[List[W.Mesh], List[str]]:
    """Returns (models, material names in .mtl order). Each model = one `o`/`g` group or one
    `usemtl` run inside a group (tobj model splitting). Faces are fan-triangulated; vertices are
    re-indexed to a single index over (v, vn) pairs."""
    text = resolver.text(filename)
    positions: List[Tuple[float, float, float]] = []
    normals: List[Tuple[float, float, float]] = []
    mtl_names: List[str] = []
    models: List[dict] = []
    cur = None
    cur_name = "unnamed_object"
    cur_mat = None

    def flush():
        nonlocal cur
        if cur is not None and cur["faces"]:
            models.append(cur)
        cur = None

    def ensure():
        nonlocal cur
        if cur is None:
            cur = {"name": cur_name, "faces": [], "mat": cur_mat}

    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        tok = line.split()
        tag = tok[0]
        if tag == "v":
            positions.append((float(tok[1]), float(tok[2]), float(tok[3])))
        elif tag == "vn":
            normals.append((float(tok[1]), float(tok[2]), float(tok[3])))
        elif tag in ("o", "g"):
            flush()
            cur_name = tok[1] if len(tok) > 1 else "unnamed_object"
        elif tag == "usemtl":
            name = tok[1]
            new_mat = mtl_names.index(name) if name in mtl_names else None
            if cur is not None and cur["faces"] and new_mat != cur["mat"]:
                flush()
            cur_mat = new_mat
            if cur is not None:
                cur["mat"] = cur_mat
        elif tag == "mtllib":
            mtl_rel = os.path.join(os.path.dirname(filename), tok[1])
            try:
                for ml in resolver.text(mtl_rel).splitlines():
                    mt = ml.split()
                    if len(mt) >= 2 and mt[0] == "newmtl":
                        mtl_names.append(mt[1])
            except LoadError:
                raise LoadError(f"Failed to load MTL file {mtl_rel} (meshes.rs:30)")
        elif tag == "f":
            ensure()
            verts = []
            for v in tok[1:]:
                parts = v.split("/")
                vi = int(parts[0])
                vi = vi - 1 if vi > 0 else len(positions) + vi
                ni = None
                if len(parts) >= 3 and parts[2] != "":
                    ni = int(parts[2])
                    ni = ni - 1 if ni > 0 else len(normals) + ni
                verts.append((vi, ni))
            for k in range(1, len(verts) - 1):
                cur["faces"].append((verts[0], verts[k], verts[k + 1]))
    flush()

    out = []
    for m in models:
        remap: Dict[Tuple[int, Optional[int]], int] = {}
        vs, ns, idx = [], [], []
        has_normals = all(n is not None for f in m["faces"] for (_, n) in f) and len(normals) > 0
        for f in m["faces"]:
            tri = []
            for key in f:
                k = key if has_normals else (key[0], None)
                if k not in remap:
                    remap[k] = len(vs)
                    vs.append(positions[k[0]])
                    if has_normals:
                        ns.append(normals[k[1]])
                tri.append(remap[k])
            idx.append(tri)
        mat = m["mat"] if m["mat"] is not None else 0
        out.append(
            W.Mesh(
                vertices=np.asarray(vs, dtype=F32).reshape(-1, 3),
                indices=np.asarray(idx, dtype=np.uint32).reshape(-1, 3),
                normals=np.asarray(ns, dtype=F32).reshape(-1, 3) if has_normals else None,
                face_material=np.full(len(idx), W.mat_pack(W.MAT_TAG_MATERIAL, mat), dtype=np.uint32),
                name=m["name"],
            )
        )
    return out, mtl_names


# ---- images ---------------------------------------------------------------------------------


def _read_hdr(path: str) -> np.ndarray:
    """Radiance RGBE (.hdr) reader: flat or new-RLE scanlines, -Y h +X w orientation."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]
        pos = end + 1
        if line == b"":
            break
    end = data.index(b"\n", pos)
    res = data[pos:end].split()
    pos = end + 1
    if res[0] != b"-Y" or res[2] != b"+X":
        raise LoadError("unsupported .hdr orientation")
    h, w = int(res[1]), int(res[3])
    rgbe = np.zeros((h, w, 4), dtype=np.uint8)
    buf = np.frombuffer(data, dtype=np.uint8)
    if len(buf) - pos == h * w * 4:
        rgbe = buf[pos:].reshape(h, w, 4).copy()
    else:
        for y in range(h):
            if not (buf[pos] == 2 and buf[pos + 1] == 2 and ((int(buf[pos + 2]) << 8) | int(buf[pos + 3])) == w):
                raise LoadError("unsupported .hdr scanline encoding")
            pos += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = int(buf[pos])
                    pos += 1
                    if n > 128:
                        n -= 128
                        rgbe[y, x : x + n, c] = buf[pos]
                        pos += 1
                    else:
                        rgbe[y, x : x + n, c] = buf[pos : pos + n]
                        pos += n
                    x += n
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e == 0, 0.0, np.ldexp(1.0, e - 136)).astype(F32)
    return (rgbe[..., :3].astype(F32) * scale[..., None]).astype(F32)


def _read_image_rgba8(path: str) -> np.ndarray:
    from PIL import Image

    return np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8)


def _read_image_luma8(path: str) -> np.ndarray:
    from PIL import Image

    return np.asarray(Image.open(path).convert("L"), dtype=np.uint8)


# ---- world construction (reference src/parsing/mod.rs:145-563) ------------------------------


class _CurveTable:
    """Interns Curve objects into World.curves and returns LUT ids."""

    def __init__(self, world: W.World, lib: Dict[str, dict], resolver: Resolver):
        self.world = world
        self.lib = lib
        self.resolver = resolver
        self.by_name: Dict[str, int] = {}
        self.curve_objs: Dict[str, C.Curve] = {}

    def add(self, curve: C.Curve, name: str) -> int:
        self.world.curves.append(curve)
        self.world.curve_names.append(name)
        return len(self.world.curves) - 1

    def resolve_obj(self, ref) -> Optional[C.Curve]:
        """CurveDataOrReference::resolve (parsing/curves.rs:380-391)."""
        if isinstance(ref, str):
            if ref in self.curve_objs:
                return self.curve_objs[ref]
            if ref not in self.lib:
                return None
            c = C.curve_from_data(self.lib[ref], self.resolver.text)
            self.curve_objs[ref] = c
            return c
        return C.curve_from_data(ref, self.resolver.text)

    def resolve(self, ref) -> Optional[int]:
        if isinstance(ref, str):
            if ref in self.by_name:
                return self.by_name[ref]
            c = self.resolve_obj(ref)
            if c is None:
                return None
            i = self.add(c, ref)
            self.by_name[ref] = i
            return i
        c = self.resolve_obj(ref)
        return self.add(c, "<literal>")


def _resolve_lib(value, resolver: Resolver) -> dict:
    return resolver.toml(value) if isinstance(value, str) else value


def construct_world(config: Config, scene_file: Optional[str] = None, resolver: Optional[Resolver] = None,
                    bake_importance_map: bool = True) -> W.World:
    resolver = resolver or Resolver()
    scene = resolver.toml(scene_file or config.scene_file)
    for k in scene:
        if k not in ("env_sampling_probability", "environment", "curves", "textures", "materials", "mediums", "meshes", "instances", "cameras"):
            raise LoadError(f"unknown scene field {k!r} (deny_unknown_fields, parsing/mod.rs:90)")
    world = W.World()
    curves_lib = _resolve_lib(scene["curves"], resolver)
    textures_lib = _resolve_lib(scene["textures"], resolver)
    materials_lib = _resolve_lib(scene["materials"], resolver)
    meshes_lib = _resolve_lib(scene["meshes"], resolver)
    ct = _CurveTable(world, curves_lib, resolver)

    # -- used sets (parsing/mod.rs:169-205)
    used_materials, used_meshes = [], []
    for inst in scene["instances"]:
        mn = inst.get("material_name")
        if mn is not None and mn not in used_materials:
            used_materials.append(mn)
        agg = inst["aggregate"]
        if agg["type"] == "Mesh" and agg["name"] not in used_meshes:
            used_meshes.append(agg["name"])

    # -- meshes (parsing/mod.rs:208-258): name or "name;index" -> Mesh ; name -> .mtl material names
    mesh_mapping: Dict[str, W.Mesh] = {}
    mesh_material_mapping: Dict[str, List[str]] = {}
    for name in used_meshes:
        if name not in meshes_lib:
            raise LoadError(f"mesh {name!r} not in mesh library")
        md = meshes_lib[name]
        models, mtl_names = load_obj_models(resolver, md["filename"])
        if md.get("mesh_index") is not None:
            mesh_mapping[name] = models[int(md["mesh_index"])]
        else:
            for i, m in enumerate(models):
                mesh_mapping[f"{name};{i}"] = m
        for mat in mtl_names:
            if mat not in used_materials:
                used_materials.append(mat)
        mesh_material_mapping[name] = list(mtl_names) if mtl_names else ["error"]

    # -- textures (parse only the used ones)
    tex_ids: Dict[str, int] = {}

    def texture_stack(name: str) -> int:
        if name in tex_ids:
            return tex_ids[name]
        if name not in textures_lib:
            raise LoadError(f"didn't find texture stack id for texture name {name!r} (material.rs:103)")
        stack = []
        for layer in textures_lib[name]:
            t = layer["type"]
            if t == "Texture1":
                cid = ct.resolve(layer["curve"])
                img = _read_image_luma8(resolver.path(layer["filename"])).astype(F32) / F32(255.0)
                tex = W.Texture(1, img[..., None], (cid, -1, -1, -1))
            elif t in ("Texture4", "SRGB"):
                if t == "SRGB":
                    base = "data/curves/basis/simple-spectral-srgb-1931.csv"
                    refs = [{"type": "TabulatedCSV", "filename": base, "column": c, "interpolation_mode": "Cubic"} for c in (1, 2, 3)]
                    refs.append({"type": "Flat", "strength": 0.0})
                else:
                    refs = layer["curves"]
                cids = tuple(ct.resolve(r) for r in refs)
                img = _read_image_rgba8(resolver.path(layer["filename"])).astype(F32) / F32(255.0)
                tex = W.Texture(4, img, cids)
            elif t in ("HDR", "EXR"):
                # parsing/texture.rs:49-120: both decode to an RGBA f32 image; alpha_fill replaces the alpha channel
                cids = tuple(ct.resolve(r) for r in layer["curves"])
                path = resolver.path(layer["filename"])
                if t == "HDR":
                    rgb = _read_hdr(path)
                else:
                    from .exr import read_exr_rgb  # uncompressed f32 scanline files (what the synthetic fixtures are written as)

                    try:
                        rgb = read_exr_rgb(path)
                    except ValueError as e:
                        raise LoadError(f"{path}: {e} (only uncompressed FLOAT scanline EXR is decoded here)")
                alpha_fill = float(layer.get("alpha_fill") or 0.0)
                alpha = np.full(rgb.shape[:2] + (1,), F32(alpha_fill), dtype=F32)
                tex = W.Texture(4, np.concatenate([rgb, alpha], axis=2), cids)
                if os.path.exists(path + ".recipe.json"):  # synthetic fixture: remember how to regenerate it (rust-pathtracer_b200/synth.py)
                    import json

                    tex.recipe = dict(json.load(open(path + ".recipe.json")), alpha_fill=alpha_fill)
            else:
                raise LoadError(f"texture type {t} unsupported")
            if any(c is None for c in tex.curves):
                raise LoadError("failed to parse curve (texture.rs:190)")
            world.textures.append(tex)
            stack.append(len(world.textures) - 1)
        world.texstacks.append(stack)
        tex_ids[name] = len(world.texstacks) - 1
        return tex_ids[name]

    # -- materials: index 0 = mauve error light (parsing/mod.rs:425-467)
    mauve_emit = ct.add(C.mauve(1.0), "<mauve>")
    mauve_bounce = ct.add(C.cie_e(0.0), "<cie_e(0)>")
    world.materials.append(W.Material(W.MATERIAL_DIFFUSE_LIGHT, "error", curve_a=mauve_bounce, curve_b=mauve_emit, sidedness=W.SIDEDNESS["Dual"]))
    world.material_names_to_ids["error"] = W.mat_pack(W.MAT_TAG_LIGHT, 0)
    for name in sorted(used_materials):
        if name not in materials_lib:
            continue
        md = materials_lib[name]
        t = md["type"]
        if t == "GGX":
            ea, eo, ka = ct.resolve(md["eta"]), ct.resolve(md["eta_o"]), ct.resolve(md["kappa"])
            if ea is None or eo is None or ka is None:
                continue  # "failed to resolve one of eta, eta_o, or kappa"
            metallic = world.curves[ka].evaluate_integral(C.BOUNDED_VISIBLE_RANGE, 100, False) > 0.0
            mat = W.Material(W.MATERIAL_GGX, name, curve_a=ea, curve_b=eo, curve_c=ka, alpha=float(md["alpha"]), metallic=metallic)
        elif t == "Lambertian":
            mat = W.Material(W.MATERIAL_LAMBERTIAN, name, texstack=texture_stack(md["texture_id"]))
        elif t in ("DiffuseLight", "SharpLight"):
            emit, bounce = ct.resolve(md["emit_color"]), ct.resolve(md["bounce_color"])
            if emit is None or bounce is None:
                continue
            mat = W.Material(
                W.MATERIAL_SHARP_LIGHT if t == "SharpLight" else W.MATERIAL_DIFFUSE_LIGHT, name,
                curve_a=bounce, curve_b=emit, sidedness=W.SIDEDNESS[md["sidedness"]],
                sharpness=(1.0 + abs(float(md["sharpness"]))) if t == "SharpLight" else 0.0,
            )
        else:
            raise LoadError(f"unknown material type {t}")
        world.materials.append(mat)
        idx = len(world.materials) - 1
        world.material_names_to_ids[name] = world.material_id(idx)

    # -- remap mesh material ids (parsing/mod.rs:472-502)
    for mname, mesh in mesh_mapping.items():
        prefix = mname.split(";")[0]
        names = mesh_material_mapping[prefix]
        remapped = np.empty_like(mesh.face_material)
        for i, m in enumerate(mesh.face_material):
            local = int(m) & 0xFFFF
            nm = names[local] if local < len(names) else "error"
            remapped[i] = world.material_names_to_ids.get(nm, W.mat_pack(W.MAT_TAG_MATERIAL, 0))
        mesh.face_material = remapped

    # -- instances (parsing/mod.rs:506-548); bundles expand in object-index order
    mesh_index: Dict[str, int] = {}
    for inst in scene["instances"]:
        agg = inst["aggregate"]
        transform = W.Transform3.from_data(inst["transform"]) if inst.get("transform") is not None else None
        mn = inst.get("material_name")
        if mn is None:
            material = W.MAT_NONE
        else:
            material = world.material_names_to_ids.get(mn, world.material_names_to_ids["error"])
        t = agg["type"]
        if t == "Mesh":
            if agg.get("index") is not None:
                raise LoadError("MeshRef with explicit index is never initialised by the reference (mesh.rs:316 panics)")
            keys = [k for k in mesh_mapping if k.startswith(agg["name"])]
            keys.sort(key=lambda k: (k.split(";")[0], int(k.split(";")[1]) if ";" in k else -1))
            for k in keys:
                if k not in mesh_index:
                    world.meshes.append(mesh_mapping[k])
                    mesh_index[k] = len(world.meshes) - 1
                world.instances.append(W.Instance(W.AGG_MESH, mesh=mesh_index[k], transform=transform, material=material))
        elif t == "Rect":
            assert agg["size"][0] > 0 and agg["size"][1] > 0
            world.instances.append(W.Instance(W.AGG_RECT, tuple(agg["origin"]), tuple(agg["size"]), W.AXIS[agg["normal"]], bool(agg["two_sided"]), transform=transform, material=material))
        elif t == "Sphere":
            assert agg["radius"] > 0
            world.instances.append(W.Instance(W.AGG_SPHERE, tuple(agg["origin"]), (agg["radius"], 0.0), transform=transform, material=material))
        elif t == "Disk":
            assert agg["radius"] > 0
            world.instances.append(W.Instance(W.AGG_DISK, tuple(agg["origin"]), (agg["radius"], 0.0), two_sided=bool(agg["two_sided"]), transform=transform, material=material))
        else:
            raise LoadError(f"unknown aggregate type {t}")

    # -- environment (parsing/environment.rs:61-181)
    env = scene["environment"]
    et = env["type"]
    mauve_curve = C.mauve(1.0)
    if et in ("Constant", "Sun"):
        cid = ct.resolve(env["color"])
        if cid is None:
            cid = ct.add(mauve_curve, "<mauve>")
        world.environment = W.Environment(kind=0 if et == "Constant" else 1, strength=float(env["strength"]), curve=cid)
        if et == "Sun":
            d = np.asarray(env["sun_direction"], dtype=F32)
            d = d / F32(np.sqrt(np.sum(d * d, dtype=F32)))
            world.environment.angular_diameter = float(env["angular_diameter"])
            world.environment.sun_direction = tuple(float(x) for x in d)
    elif et == "HDRI":
        rot = W.Transform3.from_data({"rotate": env.get("rotation")})
        try:
            ts = texture_stack(env["texture_name"])
        except LoadError:
            # "importance map texture not found, using mauve texture" (environment.rs:105-114)
            cid = ct.add(mauve_curve, "<mauve>")
            world.textures.append(W.Texture(1, np.ones((1, 1, 1), dtype=F32), (cid, -1, -1, -1)))
            world.texstacks.append([len(world.textures) - 1])
            ts = len(world.texstacks) - 1
        world.environment = W.Environment(kind=2, strength=float(env["strength"]), texstack=ts, rotation=rot)
        im = env.get("importance_map")
        if im is not None and float(env["strength"]) > 0.0:
            lum = C.curve_from_data(im["luminance_curve"], resolver.text) if im.get("luminance_curve") else C.y_bar_curve()
            world.environment.imap_request = (int(im["height"]), int(im["width"]), lum)
            if bake_importance_map:  # host bake; with False the request is left for CudaRenderer to bake on the device
                from .importance_map import bake_importance_map as _bake

                _bake(world, int(im["height"]), int(im["width"]), lum, C.BOUNDED_VISIBLE_RANGE)
    else:
        raise LoadError(f"unknown environment type {et}")

    # -- cameras (parsing/cameras.rs:116-204): one aspect-corrected camera per render setting
    by_name = {}
    for cam in scene["cameras"]:
        if cam["type"] not in ("SimpleCamera", "PanoramaCamera"):
            continue  # RealisticCamera sits behind a cargo feature (parsing/cameras.rs:101-102): out of scope
        v_up = np.asarray(cam.get("v_up") or [0.0, 0.0, 1.0], dtype=F32)
        v_up = v_up / F32(np.sqrt(np.sum(v_up * v_up, dtype=F32)))
        if cam["type"] == "PanoramaCamera":  # parsing/cameras.rs:150-160
            by_name[cam["name"]] = W.Camera.new_panorama(cam["name"], cam["look_from"], cam["look_at"], v_up, cam["fov"][0], cam["fov"][1])
            continue
        by_name[cam["name"]] = W.Camera.new(
            cam["name"], cam["look_from"], cam["look_at"], v_up, cam["vfov"],
            cam.get("focal_distance") if cam.get("focal_distance") is not None else 10.0,
            cam.get("aperture_diameter") if cam.get("aperture_diameter") is not None else 0.01,
        )
    for rs in config.render_settings:
        if rs.camera_id not in by_name:
            raise LoadError(f"camera {rs.camera_id!r} is not defined by the scene (reference panics at parsing/cameras.rs:196)")
        config.camera_names_to_index[rs.camera_id] = len(world.cameras)
        world.camera_names_to_index[rs.camera_id] = len(world.cameras)
        world.cameras.append(by_name[rs.camera_id].with_aspect_ratio(rs.width / rs.height))

    esp = scene.get("env_sampling_probability")
    world.env_sampling_probability = float(esp) if esp is not None else 0.5
    world.compute_lights()
    return world
