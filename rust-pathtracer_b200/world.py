"""Host-side `World`: the Python mirror of reference src/world/mod.rs:18-28 and the geometry /
material / camera types it owns. Pure data + the small amount of construction math the reference
does at parse time (Transform3 stacks, camera frame). Nothing here runs per ray.

`math::Transform3` is external and unpinned: `forward` maps local->world,
`reverse` = inverse; `from_stack(scale, rotate, translate)` composes T*R*S.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import curves as C

F32 = np.float32

MAT_TAG_MATERIAL = 0
MAT_TAG_LIGHT = 1
MAT_NONE = 0xFFFFFFFF


def mat_pack(tag: int, index: int) -> int:
    return ((tag & 0xFF) << 16) | (index & 0xFFFF)


# ---- Transform3 -----------------------------------------------------------------------------


def _eye() -> np.ndarray:
    return np.eye(4, dtype=np.float64)


@dataclass
class Transform3:
    forward: np.ndarray  # 4x4 float64 (rounded to f32 when flattened)
    reverse: np.ndarray

    @staticmethod
    def from_matrix(m: np.ndarray) -> "Transform3":
        return Transform3(m, np.linalg.inv(m))

    @staticmethod
    def from_scale(s) -> "Transform3":
        m = _eye()
        m[0, 0], m[1, 1], m[2, 2] = s
        return Transform3.from_matrix(m)

    @staticmethod
    def from_translation(t) -> "Transform3":
        m = _eye()
        m[:3, 3] = t
        return Transform3.from_matrix(m)

    @staticmethod
    def from_axis_angle(axis, radians: float) -> "Transform3":
        x, y, z = axis
        c, s = math.cos(radians), math.sin(radians)
        k = 1.0 - c
        m = _eye()
        m[:3, :3] = [
            [c + x * x * k, x * y * k - z * s, x * z * k + y * s],
            [y * x * k + z * s, c + y * y * k, y * z * k - x * s],
            [z * x * k - y * s, z * y * k + x * s, c + z * z * k],
        ]
        return Transform3.from_matrix(m)

    def __mul__(self, other: "Transform3") -> "Transform3":
        return Transform3(self.forward @ other.forward, other.reverse @ self.reverse)

    @staticmethod
    def from_stack(scale, rotate, translate) -> "Transform3":
        stack = [t for t in (scale, rotate, translate) if t is not None]
        out = Transform3(_eye(), _eye())
        for t in stack:
            out = t * out
        return out

    @staticmethod
    def from_data(data: dict) -> "Transform3":
        """Transform3Data -> Transform3 (reference src/parsing/instance.rs:40-71)."""
        scale = Transform3.from_scale(data["scale"]) if data.get("scale") is not None else None
        rotate = None
        for rot in data.get("rotate") or []:
            ax = np.asarray(rot["axis"], dtype=np.float64)
            ax = ax / np.linalg.norm(ax)
            t = Transform3.from_axis_angle(ax, math.pi * rot["angle"] / 180.0)
            rotate = t if rotate is None else t * rotate
        translate = Transform3.from_translation(data["translate"]) if data.get("translate") is not None else None
        return Transform3.from_stack(scale, rotate, translate)


# ---- geometry -------------------------------------------------------------------------------

AGG_RECT, AGG_SPHERE, AGG_DISK, AGG_MESH = 0, 1, 2, 3
AXIS = {"X": 0, "Y": 1, "Z": 2}


@dataclass
class Mesh:
    """reference src/geometry/mesh.rs:257-268 (triangles only; quads are fan-triangulated by the loader)."""

    vertices: np.ndarray  # (nv,3) f32
    indices: np.ndarray  # (nf,3) u32
    normals: Optional[np.ndarray]  # (nv,3) f32 or None
    face_material: np.ndarray  # (nf,) u32 packed MaterialId (local .mtl index until remapped)
    name: str = ""


@dataclass
class Instance:
    """reference src/geometry/instance.rs:9-15."""

    kind: int
    origin: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    size: Tuple[float, float] = (0.0, 0.0)
    axis: int = 2
    two_sided: bool = False
    mesh: int = -1
    transform: Optional[Transform3] = None
    material: int = MAT_NONE


# ---- materials ------------------------------------------------------------------------------

MATERIAL_LAMBERTIAN, MATERIAL_GGX, MATERIAL_DIFFUSE_LIGHT, MATERIAL_SHARP_LIGHT = 0, 1, 2, 3
SIDEDNESS = {"Forward": 0, "Reverse": 1, "Dual": 2}


@dataclass
class Material:
    type: int
    name: str = ""
    texstack: int = -1
    curve_a: int = -1
    curve_b: int = -1
    curve_c: int = -1
    alpha: float = 0.0
    sharpness: float = 0.0
    sidedness: int = 0
    metallic: bool = False

    @property
    def is_light(self) -> bool:
        return self.type in (MATERIAL_DIFFUSE_LIGHT, MATERIAL_SHARP_LIGHT)


@dataclass
class Texture:
    channels: int
    texels: np.ndarray  # (h, w, channels) f32
    curves: Tuple[int, int, int, int]
    recipe: Optional[dict] = None  # synth.py recipe that regenerates `texels` (synthetic environment maps; scene blobs store it instead)


@dataclass
class Environment:
    kind: int = 0  # 0 Constant, 1 Sun, 2 HDR
    strength: float = 0.0
    curve: int = -1
    angular_diameter: float = 0.0
    sun_direction: Tuple[float, float, float] = (0.0, 0.0, 1.0)
    texstack: int = -1
    rotation: Optional[Transform3] = None
    imap_row_pdf: Optional[np.ndarray] = None  # (rows, cols)
    imap_row_cdf: Optional[np.ndarray] = None
    imap_marginal_pdf: Optional[np.ndarray] = None
    imap_marginal_cdf: Optional[np.ndarray] = None
    imap_marginal_integral: float = 1.0
    # ImportanceMap::Unbaked request of the scene file: (vertical_resolution, horizontal_resolution, luminance Curve)
    imap_request: Optional[tuple] = None


@dataclass
class Camera:
    """ProjectiveCamera constants (reference src/camera/projective_camera.rs:27-94,121-133); kind 1 = PanoramaCamera
    (src/camera/panorama_camera.rs:18-62): origin, frame (u, v, w = +direction) and the angle spans in radians."""

    name: str
    origin: np.ndarray
    u: np.ndarray
    v: np.ndarray
    w: np.ndarray
    lower_left: np.ndarray
    horizontal: np.ndarray
    vertical: np.ndarray
    aperture_diameter: float
    vfov: float
    focal_distance: float
    kind: int = 0
    angle_span: Tuple[float, float] = (0.0, 0.0)

    @staticmethod
    def new_panorama(name, look_from, look_at, v_up, horizontal_fov, vertical_fov) -> "Camera":
        """PanoramaCamera::new (panorama_camera.rs:18-62); fovs in degrees, clamped to (2 pi, pi)."""
        f = lambda a: np.asarray(a, dtype=F32)
        look_from, look_at, v_up = f(look_from), f(look_at), f(v_up)

        def normalized(a):
            return (a / F32(np.sqrt(np.sum(a * a, dtype=F32)))).astype(F32)

        w = normalized(look_at - look_from)
        u = normalized(np.cross(v_up, w).astype(F32))
        v = normalized(np.cross(w, u).astype(F32))
        hf = float(np.clip(F32(np.deg2rad(F32(horizontal_fov))), F32(0.0), F32(2.0 * np.pi)))
        vf = float(np.clip(F32(np.deg2rad(F32(vertical_fov))), F32(0.0), F32(np.pi)))
        z = f([0, 0, 0])
        return Camera(name, look_from, u, v, w, z, z, z, 0.0, 0.0, 0.0, kind=1, angle_span=(hf, vf))

    @staticmethod
    def new(name, look_from, look_at, v_up, vfov, focal_distance, aperture_diameter) -> "Camera":
        f = lambda a: np.asarray(a, dtype=F32)
        look_from, look_at, v_up = f(look_from), f(look_at), f(v_up)

        def normalized(a):
            return (a / F32(np.sqrt(np.sum(a * a, dtype=F32)))).astype(F32)

        direction = normalized(look_at - look_from)
        w = -direction
        u = -normalized(np.cross(v_up, w).astype(F32))
        v = normalized(np.cross(w, u).astype(F32))
        cam = Camera(name, look_from, u, v, w, f([0, 0, 0]), f([0, 0, 0]), f([0, 0, 0]), float(aperture_diameter), float(vfov), float(focal_distance))
        return cam.with_aspect_ratio(1.0)

    def with_aspect_ratio(self, aspect: float) -> "Camera":
        if self.kind == 1:
            return self  # panorama_camera.rs:92-94
        theta = F32(np.deg2rad(F32(self.vfov)))
        half_height = F32(np.tan(theta / F32(2.0)))
        half_width = F32(aspect) * half_height
        fd = F32(self.focal_distance)
        ll = self.origin - self.u * half_width * fd - self.v * half_height * fd - self.w * fd
        hor = self.u * F32(2.0) * half_width * fd
        ver = self.v * F32(2.0) * half_height * fd
        return Camera(self.name, self.origin, self.u, self.v, self.w, ll.astype(F32), hor.astype(F32), ver.astype(F32), self.aperture_diameter, self.vfov, self.focal_distance)


@dataclass
class World:
    """Flattenable scene. Index conventions follow reference src/parsing/mod.rs:440-467:
    materials[0] is the mauve error light; lights get MaterialId::Light(i)."""

    instances: List[Instance] = field(default_factory=list)
    meshes: List[Mesh] = field(default_factory=list)
    lights: List[int] = field(default_factory=list)
    materials: List[Material] = field(default_factory=list)
    curves: List[C.Curve] = field(default_factory=list)  # evaluated into LUTs at flatten time
    curve_names: List[str] = field(default_factory=list)
    textures: List[Texture] = field(default_factory=list)
    texstacks: List[List[int]] = field(default_factory=list)
    environment: Environment = field(default_factory=Environment)
    env_sampling_probability: float = 0.5
    cameras: List[Camera] = field(default_factory=list)
    camera_names_to_index: Dict[str, int] = field(default_factory=dict)
    material_names_to_ids: Dict[str, int] = field(default_factory=dict)

    def material_id(self, index: int) -> int:
        m = self.materials[index]
        return mat_pack(MAT_TAG_LIGHT if m.is_light else MAT_TAG_MATERIAL, index)

    def compute_lights(self) -> None:
        """reference src/world/mod.rs:42-66: one entry per light instance; for meshes one entry
        per light-material triangle."""
        self.lights = []
        for iid, inst in enumerate(self.instances):
            if inst.kind == AGG_MESH:
                mesh = self.meshes[inst.mesh]
                for m in mesh.face_material:
                    if (int(m) >> 16) & 0xFF == MAT_TAG_LIGHT:
                        self.lights.append(iid)
            else:
                m = inst.material if inst.material != MAT_NONE else mat_pack(MAT_TAG_MATERIAL, 0)
                if (m >> 16) & 0xFF == MAT_TAG_LIGHT:
                    self.lights.append(iid)
