"""ctypes binding of include/rpt.h + World -> RptSceneDesc flattening + scene blobs.

This is the reference-side binding a maintainer would write in Rust (`extern "C"` block, see
INTEGRATION.md); here it is ctypes because the image has no Rust toolchain. The product library
(`librpt_b200.so`, CUDA) and the CPU oracle (`oracle/_build/librpt_oracle.so`, tests only) take
the same structs.
"""
from __future__ import annotations

import ctypes as ct
import os
from typing import List, Optional

import numpy as np

from . import curves as C
from . import world as W

F32 = np.float32
ABI_VERSION = 8
PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "librpt_b200.so")
ORACLE_LIB_PATH = os.path.join(REPO_ROOT, "oracle", "_build", "librpt_oracle.so")

c_f = ct.c_float
c_u32 = ct.c_uint32
c_i32 = ct.c_int32
c_u64 = ct.c_uint64
PF = ct.POINTER(c_f)
PU = ct.POINTER(c_u32)


class RptInstance(ct.Structure):
    _fields_ = [
        ("kind", c_u32), ("origin", c_f * 3), ("size", c_f * 2), ("axis", c_u32), ("two_sided", c_u32),
        ("mesh", c_i32), ("has_transform", c_u32), ("forward", c_f * 16), ("reverse", c_f * 16), ("material", c_u32),
    ]


class RptMesh(ct.Structure):
    _fields_ = [
        ("num_vertices", c_u32), ("num_faces", c_u32), ("vertices", PF), ("indices", PU), ("normals", PF), ("face_material", PU),
    ]


class RptMaterial(ct.Structure):
    _fields_ = [
        ("type", c_u32), ("texstack", c_i32), ("curve_a", c_i32), ("curve_b", c_i32), ("curve_c", c_i32),
        ("alpha", c_f), ("sharpness", c_f), ("sidedness", c_u32), ("metallic", c_u32),
    ]


class RptTexture(ct.Structure):
    _fields_ = [("channels", c_u32), ("width", c_u32), ("height", c_u32), ("texels", PF), ("curves", c_i32 * 4)]


class RptTexStack(ct.Structure):
    _fields_ = [("first", c_u32), ("count", c_u32)]


class RptEnvironment(ct.Structure):
    _fields_ = [
        ("kind", c_u32), ("strength", c_f), ("curve", c_i32), ("angular_diameter", c_f), ("sun_direction", c_f * 3),
        ("texstack", c_i32), ("rot_forward", c_f * 16), ("rot_reverse", c_f * 16),
        ("imap_rows", c_u32), ("imap_cols", c_u32), ("imap_row_pdf", PF), ("imap_row_cdf", PF),
        ("imap_marginal_n", c_u32), ("imap_marginal_pdf", PF), ("imap_marginal_cdf", PF), ("imap_marginal_integral", c_f),
    ]


class RptCamera(ct.Structure):
    _fields_ = [
        ("origin", c_f * 3), ("u", c_f * 3), ("v", c_f * 3), ("w", c_f * 3), ("lower_left", c_f * 3),
        ("horizontal", c_f * 3), ("vertical", c_f * 3), ("aperture_diameter", c_f), ("kind", c_u32), ("angle_span", c_f * 2),
    ]


class RptSceneDesc(ct.Structure):
    _fields_ = [
        ("abi_version", c_u32),
        ("num_instances", c_u32), ("instances", ct.POINTER(RptInstance)),
        ("num_meshes", c_u32), ("meshes", ct.POINTER(RptMesh)),
        ("num_lights", c_u32), ("lights", PU),
        ("num_materials", c_u32), ("materials", ct.POINTER(RptMaterial)),
        ("num_curves", c_u32), ("num_lambda", c_u32), ("lut_lambda_lo", c_f), ("lut_lambda_hi", c_f),
        ("curve_lut", PF), ("cie_lut", PF),
        ("num_textures", c_u32), ("textures", ct.POINTER(RptTexture)),
        ("num_texstack_textures", c_u32), ("texstack_textures", PU),
        ("num_texstacks", c_u32), ("texstacks", ct.POINTER(RptTexStack)),
        ("environment", RptEnvironment),
        ("env_sampling_probability", c_f),
        ("num_cameras", c_u32), ("cameras", ct.POINTER(RptCamera)),
    ]


class RptRenderParams(ct.Structure):
    _fields_ = [
        ("width", c_u32), ("height", c_u32), ("spp", c_u32), ("spp_offset", c_u32), ("spp_total", c_u32),
        ("min_bounces", c_u32), ("max_bounces", c_u32), ("light_samples", c_u32), ("only_direct", c_u32),
        ("lambda_lo", c_f), ("lambda_hi", c_f), ("camera", c_u32), ("seed", c_u64), ("flags", c_u32), ("reserved", c_u32),
    ]


FLAG_KERNEL_TIMES = 1  # RPT_FLAG_KERNEL_TIMES: CUDA events around every kernel launch -> Scene.kernel_times()
FLAG_BVH_STATS = 2     # RPT_FLAG_BVH_STATS: RptCounters.walk_* / shadow_* are filled


class RptCounters(ct.Structure):
    _fields_ = [
        ("camera_rays", c_u64), ("bounce_rays", c_u64), ("shadow_rays", c_u64), ("light_rays", c_u64),
        ("env_hits", c_u64), ("segments", c_u64), ("true_rays", c_u64), ("kernel_launches", c_u64),
        ("shadow_rays_traced", c_u64), ("walk_nodes", c_u64), ("walk_tris", c_u64), ("walk_insts", c_u64),
        ("shadow_nodes", c_u64), ("shadow_tris", c_u64), ("shadow_insts", c_u64), ("nee_vertices", c_u64), ("device_ms", ct.c_double),
    ]

    def as_dict(self) -> dict:
        return {k: (float(getattr(self, k)) if k == "device_ms" else int(getattr(self, k))) for k, _ in self._fields_}


class RptOutputSettings(ct.Structure):
    _fields_ = [
        ("tonemapper", c_u32), ("luminance_only", c_u32), ("exposure", c_f), ("key_value", c_f), ("white_point", c_f),
        ("colorspace", c_u32), ("factor", c_f),
    ]


class RptImapBake(ct.Structure):
    _fields_ = [
        ("rows", c_u32), ("cols", c_u32), ("num_samples", c_u32), ("lambda_lo", c_f), ("lambda_hi", c_f),
        ("luminance", PF), ("basis", PF),
    ]


class RptKernelTime(ct.Structure):
    _fields_ = [("name", ct.c_char_p), ("launches", c_u32), ("ms", c_f)]


class RptSceneStats(ct.Structure):
    _fields_ = [
        ("tlas_nodes", c_u64), ("blas_nodes", c_u64), ("triangles", c_u64), ("instances", c_u64),
        ("node_bytes", c_u64), ("triangle_bytes", c_u64), ("scene_bytes_total", c_u64),
    ]


class RptMultiTimes(ct.Structure):
    _fields_ = [
        ("method", c_u32), ("devices", c_u32), ("render_device_ms_max", ct.c_double), ("exchange_device_ms", ct.c_double),
        ("render_wall_ms", ct.c_double), ("exchange_wall_ms", ct.c_double), ("download_wall_ms", ct.c_double),
    ]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["method"] = {0: "peer", 1: "nccl"}[int(self.method)]
        return d


# Every symbol include/rpt.h declares (tests/test_abi.py checks the .so exports all of them).
RPT_SYMBOLS = [
    "rpt_last_error", "rpt_abi_version", "rpt_device_count", "rpt_scene_create", "rpt_scene_destroy",
    "rpt_render_pt", "rpt_render_pt_device", "rpt_trace_primary", "rpt_trace_rays", "rpt_film_scale",
    "rpt_last_kernel_times", "rpt_scene_stats", "rpt_output_film", "rpt_scene_bake_importance_map", "rpt_probe_bandwidth",
    "rpt_multi_create", "rpt_multi_destroy", "rpt_multi_scene", "rpt_multi_bake_importance_map", "rpt_multi_render_pt", "rpt_render_pt_multi",
]


class RptError(RuntimeError):
    pass


def _declare(lib: ct.CDLL, prefix: str = "rpt") -> ct.CDLL:
    p = prefix
    getattr(lib, f"{p}_last_error").restype = ct.c_char_p
    getattr(lib, f"{p}_abi_version").restype = c_u32
    f = getattr(lib, f"{p}_scene_create")
    f.argtypes = [ct.POINTER(RptSceneDesc), ct.c_int, ct.POINTER(ct.c_void_p)]
    f.restype = ct.c_int
    f = getattr(lib, f"{p}_scene_destroy")
    f.argtypes = [ct.c_void_p]
    f.restype = ct.c_int
    f = getattr(lib, f"{p}_render_pt")
    f.argtypes = [ct.c_void_p, ct.POINTER(RptRenderParams), ct.c_void_p, ct.POINTER(RptCounters)]
    f.restype = ct.c_int
    f = getattr(lib, f"{p}_trace_primary")
    f.argtypes = [ct.c_void_p, ct.POINTER(RptRenderParams), ct.c_void_p, ct.c_void_p, ct.c_void_p]
    f.restype = ct.c_int
    f = getattr(lib, f"{p}_trace_rays")
    f.argtypes = [ct.c_void_p, c_u32, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p]
    f.restype = ct.c_int
    return lib


_LIB: Optional[ct.CDLL] = None


def load_library(path: Optional[str] = None) -> ct.CDLL:
    """Loads the CUDA product library. Fails loudly: there is no CPU fallback."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RptError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                       "There is no CPU fallback for the product path.")
    lib = _declare(ct.CDLL(p))
    lib.rpt_device_count.argtypes = [ct.POINTER(ct.c_int)]
    lib.rpt_device_count.restype = ct.c_int
    lib.rpt_render_pt_device.argtypes = [ct.c_void_p, ct.POINTER(RptRenderParams), ct.POINTER(ct.c_void_p), ct.POINTER(RptCounters)]
    lib.rpt_render_pt_device.restype = ct.c_int
    lib.rpt_film_scale.argtypes = [ct.c_void_p, ct.c_void_p, c_u64, c_f]
    lib.rpt_film_scale.restype = ct.c_int
    lib.rpt_last_kernel_times.argtypes = [ct.c_void_p, ct.POINTER(RptKernelTime), c_u32, PU]
    lib.rpt_last_kernel_times.restype = ct.c_int
    lib.rpt_scene_stats.argtypes = [ct.c_void_p, ct.POINTER(RptSceneStats)]
    lib.rpt_scene_stats.restype = ct.c_int
    lib.rpt_output_film.argtypes = [ct.c_void_p, ct.c_void_p, c_u32, c_u32, ct.POINTER(RptOutputSettings), ct.c_void_p, ct.c_void_p, ct.c_void_p]
    lib.rpt_output_film.restype = ct.c_int
    lib.rpt_multi_create.argtypes = [ct.POINTER(RptSceneDesc), ct.POINTER(ct.c_int), ct.c_int, ct.POINTER(ct.c_void_p)]
    lib.rpt_multi_create.restype = ct.c_int
    lib.rpt_multi_destroy.argtypes = [ct.c_void_p]
    lib.rpt_multi_destroy.restype = ct.c_int
    lib.rpt_multi_scene.argtypes = [ct.c_void_p, ct.c_int, ct.POINTER(ct.c_void_p)]
    lib.rpt_multi_scene.restype = ct.c_int
    lib.rpt_multi_bake_importance_map.argtypes = [ct.c_void_p, ct.POINTER(RptImapBake)]
    lib.rpt_multi_bake_importance_map.restype = ct.c_int
    lib.rpt_multi_render_pt.argtypes = [ct.c_void_p, ct.POINTER(RptRenderParams), ct.c_void_p, ct.POINTER(RptCounters), ct.POINTER(RptMultiTimes)]
    lib.rpt_multi_render_pt.restype = ct.c_int
    lib.rpt_render_pt_multi.argtypes = [ct.POINTER(RptSceneDesc), ct.POINTER(ct.c_int), ct.c_int, ct.POINTER(RptRenderParams), ct.c_void_p, ct.POINTER(RptCounters)]
    lib.rpt_render_pt_multi.restype = ct.c_int
    lib.rpt_probe_bandwidth.argtypes = [ct.c_int, c_u64, c_u32, ct.c_int, ct.POINTER(ct.c_double)]
    lib.rpt_probe_bandwidth.restype = ct.c_int
    if lib.rpt_abi_version() != ABI_VERSION:
        raise RptError(f"ABI mismatch: library {lib.rpt_abi_version()} vs binding {ABI_VERSION}; rebuild")
    if path is None:
        _LIB = lib
    return lib


def _arr(values, ctype, n):
    a = (ctype * n)()
    for i, v in enumerate(values):
        a[i] = v
    return a


def _mat16(m: np.ndarray):
    return _arr(np.asarray(m, dtype=F32).reshape(16).tolist(), c_f, 16)


class FlatScene:
    """Owns the numpy arrays + ctypes structs an RptSceneDesc points into."""

    def __init__(self, world: W.World, lambda_lo: float, lambda_hi: float, num_lambda: int = 1024):
        self.keep: List[object] = []
        self.world = world
        self.lambda_bounds = (float(lambda_lo), float(lambda_hi))
        k = self.keep
        d = RptSceneDesc()
        d.abi_version = ABI_VERSION

        def fptr(a: np.ndarray):
            a = np.ascontiguousarray(a, dtype=F32)
            k.append(a)
            return a.ctypes.data_as(PF)

        def uptr(a: np.ndarray):
            a = np.ascontiguousarray(a, dtype=np.uint32)
            k.append(a)
            return a.ctypes.data_as(PU)

        # meshes
        meshes = (RptMesh * max(1, len(world.meshes)))()
        for i, m in enumerate(world.meshes):
            meshes[i].num_vertices = len(m.vertices)
            meshes[i].num_faces = len(m.indices)
            meshes[i].vertices = fptr(m.vertices)
            meshes[i].indices = uptr(m.indices)
            meshes[i].normals = fptr(m.normals) if m.normals is not None else ct.cast(None, PF)
            meshes[i].face_material = uptr(m.face_material)
        k.append(meshes)
        d.num_meshes, d.meshes = len(world.meshes), meshes

        # instances
        insts = (RptInstance * max(1, len(world.instances)))()
        for i, s in enumerate(world.instances):
            r = insts[i]
            r.kind = s.kind
            r.origin = _arr(s.origin, c_f, 3)
            r.size = _arr(s.size, c_f, 2)
            r.axis, r.two_sided, r.mesh = s.axis, int(s.two_sided), s.mesh
            r.has_transform = int(s.transform is not None)
            fwd = s.transform.forward if s.transform is not None else np.eye(4)
            rev = s.transform.reverse if s.transform is not None else np.eye(4)
            r.forward, r.reverse = _mat16(fwd), _mat16(rev)
            r.material = s.material
        k.append(insts)
        d.num_instances, d.instances = len(world.instances), insts

        d.num_lights, d.lights = len(world.lights), uptr(np.asarray(world.lights, dtype=np.uint32))

        mats = (RptMaterial * max(1, len(world.materials)))()
        for i, m in enumerate(world.materials):
            r = mats[i]
            r.type, r.texstack = m.type, m.texstack
            r.curve_a, r.curve_b, r.curve_c = m.curve_a, m.curve_b, m.curve_c
            r.alpha, r.sharpness, r.sidedness, r.metallic = m.alpha, m.sharpness, m.sidedness, int(m.metallic)
        k.append(mats)
        d.num_materials, d.materials = len(world.materials), mats

        # curve LUTs (evaluate() on the uniform grid; include/rpt.h `curve_lut`)
        grid = C.lut_grid(lambda_lo, lambda_hi, num_lambda)
        self.grid = grid
        lut = np.zeros((max(1, len(world.curves)), num_lambda), dtype=F32)
        for i, c in enumerate(world.curves):
            lut[i] = c.evaluate_power(grid)
        self.curve_lut = lut
        d.num_curves, d.num_lambda = len(world.curves), num_lambda
        d.lut_lambda_lo, d.lut_lambda_hi = lambda_lo, lambda_hi
        d.curve_lut = fptr(lut)
        self.cie_lut = C.cie_xyz_bar(grid)
        d.cie_lut = fptr(self.cie_lut)

        texs = (RptTexture * max(1, len(world.textures)))()
        for i, t in enumerate(world.textures):
            texs[i].channels = t.channels
            texs[i].height, texs[i].width = t.texels.shape[0], t.texels.shape[1]
            texs[i].texels = fptr(t.texels)
            texs[i].curves = _arr(t.curves, c_i32, 4)
        k.append(texs)
        d.num_textures, d.textures = len(world.textures), texs
        flat, stacks = [], (RptTexStack * max(1, len(world.texstacks)))()
        for i, s in enumerate(world.texstacks):
            stacks[i].first, stacks[i].count = len(flat), len(s)
            flat.extend(s)
        k.append(stacks)
        d.num_texstack_textures, d.texstack_textures = len(flat), uptr(np.asarray(flat, dtype=np.uint32))
        d.num_texstacks, d.texstacks = len(world.texstacks), stacks

        e, we = d.environment, world.environment
        e.kind, e.strength, e.curve = we.kind, we.strength, we.curve
        e.angular_diameter = we.angular_diameter
        e.sun_direction = _arr(we.sun_direction, c_f, 3)
        e.texstack = we.texstack
        rot = we.rotation
        e.rot_forward = _mat16(rot.forward if rot is not None else np.eye(4))
        e.rot_reverse = _mat16(rot.reverse if rot is not None else np.eye(4))
        if we.imap_row_pdf is not None:
            e.imap_rows, e.imap_cols = we.imap_row_pdf.shape
            e.imap_row_pdf, e.imap_row_cdf = fptr(we.imap_row_pdf), fptr(we.imap_row_cdf)
            e.imap_marginal_n = len(we.imap_marginal_cdf)
            e.imap_marginal_pdf, e.imap_marginal_cdf = fptr(we.imap_marginal_pdf), fptr(we.imap_marginal_cdf)
            e.imap_marginal_integral = we.imap_marginal_integral
        d.env_sampling_probability = world.env_sampling_probability

        cams = (RptCamera * max(1, len(world.cameras)))()
        for i, c in enumerate(world.cameras):
            r = cams[i]
            for name in ("origin", "u", "v", "w", "lower_left", "horizontal", "vertical"):
                setattr(r, name, _arr(np.asarray(getattr(c, name), dtype=F32).tolist(), c_f, 3))
            r.aperture_diameter = c.aperture_diameter
            r.kind = c.kind
            r.angle_span = _arr([float(c.angle_span[0]), float(c.angle_span[1])], c_f, 2)
        k.append(cams)
        d.num_cameras, d.cameras = len(world.cameras), cams
        self.desc = d


def probe_bandwidth(lib: ct.CDLL, device: int, nbytes: int, reps: int, mode: int) -> float:
    """rpt_probe_bandwidth -> GB/s (mode 0 streaming read, 1 copy, 2 random 64-byte gathers)."""
    out = ct.c_double()
    if lib.rpt_probe_bandwidth(device, nbytes, reps, mode, ct.byref(out)) != 0:
        raise RptError(lib.rpt_last_error().decode("utf-8", "replace"))
    return float(out.value)


def debug_env_roundtrip(lib: ct.CDLL, device: int, u: np.ndarray, v: np.ndarray, fast: bool):
    """rpt_debug_env_roundtrip (a test hook of the product library, not part of rpt.h): the uv -> direction -> uv round trip of an
    unrotated HDR environment on the device, through libm (fast=False) or the libm-free path (fast=True)."""
    u = np.ascontiguousarray(u, np.float32)
    v = np.ascontiguousarray(v, np.float32)
    uo, vo = np.empty_like(u), np.empty_like(v)
    fp = ct.POINTER(ct.c_float)
    fn = lib.rpt_debug_env_roundtrip
    fn.argtypes = [ct.c_int, c_u32, fp, fp, ct.c_int, fp, fp]
    fn.restype = ct.c_int
    if fn(device, len(u), u.ctypes.data_as(fp), v.ctypes.data_as(fp), int(fast), uo.ctypes.data_as(fp), vo.ctypes.data_as(fp)) != 0:
        raise RptError(lib.rpt_last_error().decode("utf-8", "replace"))
    return uo, vo


class Scene:
    """RAII wrapper over an RptScene* of either library (product or oracle)."""

    def __init__(self, lib: ct.CDLL, flat: FlatScene, device: int = 0, prefix: str = "rpt"):
        self.lib, self.flat, self.prefix = lib, flat, prefix
        self.handle = ct.c_void_p()
        rc = getattr(lib, f"{prefix}_scene_create")(ct.byref(flat.desc), device, ct.byref(self.handle))
        if rc != 0:
            raise RptError(self._err())

    def _err(self) -> str:
        return getattr(self.lib, f"{self.prefix}_last_error")().decode("utf-8", "replace")

    def _fn(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def render_pt(self, params: RptRenderParams):
        film = np.zeros((params.height, params.width, 4), dtype=F32)
        counters = RptCounters()
        rc = self._fn("render_pt")(self.handle, ct.byref(params), film.ctypes.data_as(ct.c_void_p), ct.byref(counters))
        if rc != 0:
            raise RptError(self._err())
        return film, counters

    def render_pt_into(self, params: RptRenderParams, film_ptr: int) -> RptCounters:
        counters = RptCounters()
        rc = self._fn("render_pt")(self.handle, ct.byref(params), ct.c_void_p(film_ptr), ct.byref(counters))
        if rc != 0:
            raise RptError(self._err())
        return counters

    def render_pt_device(self, params: RptRenderParams):
        ptr = ct.c_void_p()
        counters = RptCounters()
        rc = self.lib.rpt_render_pt_device(self.handle, ct.byref(params), ct.byref(ptr), ct.byref(counters))
        if rc != 0:
            raise RptError(self._err())
        return ptr.value, counters

    def film_scale(self, film_dev: int, n_float4: int, scale: float) -> None:
        if self.lib.rpt_film_scale(self.handle, ct.c_void_p(film_dev), n_float4, scale) != 0:
            raise RptError(self._err())

    def trace_primary(self, params: RptRenderParams):
        n = params.width * params.height
        inst = np.zeros(n, dtype=np.uint32)
        prim = np.zeros(n, dtype=np.uint32)
        t = np.zeros(n, dtype=F32)
        rc = self._fn("trace_primary")(self.handle, ct.byref(params), inst.ctypes.data_as(ct.c_void_p), prim.ctypes.data_as(ct.c_void_p), t.ctypes.data_as(ct.c_void_p))
        if rc != 0:
            raise RptError(self._err())
        return inst, prim, t

    def trace_rays(self, origins: np.ndarray, dirs: np.ndarray, tmax: np.ndarray):
        o = np.ascontiguousarray(origins, dtype=F32)
        d = np.ascontiguousarray(dirs, dtype=F32)
        tm = np.ascontiguousarray(tmax, dtype=F32)
        n = len(tm)
        inst = np.zeros(n, dtype=np.uint32)
        prim = np.zeros(n, dtype=np.uint32)
        t = np.zeros(n, dtype=F32)
        vp = lambda a: a.ctypes.data_as(ct.c_void_p)
        rc = self._fn("trace_rays")(self.handle, n, vp(o), vp(d), vp(tm), vp(inst), vp(prim), vp(t))
        if rc != 0:
            raise RptError(self._err())
        return inst, prim, t

    def output_film(self, settings: "RptOutputSettings", film: Optional[np.ndarray], width: int, height: int):
        """-> (rgb_linear (H, W, 3) f32, rgba8 (H, W, 4) u8, l_w (4,) f32). film=None uses the device film of the last render."""
        rgb = np.zeros((height, width, 3), dtype=F32)
        rgba = np.zeros((height, width, 4), dtype=np.uint8)
        lw = np.zeros(4, dtype=F32)
        fp = None
        if film is not None:
            film = np.ascontiguousarray(film, dtype=F32)
            fp = film.ctypes.data_as(ct.c_void_p)
        vp = lambda a: a.ctypes.data_as(ct.c_void_p)
        fn = self._fn("output_film")
        fn.argtypes = [ct.c_void_p, ct.c_void_p, c_u32, c_u32, ct.POINTER(RptOutputSettings), ct.c_void_p, ct.c_void_p, ct.c_void_p]
        fn.restype = ct.c_int
        if fn(self.handle, fp, width, height, ct.byref(settings), vp(rgb), vp(rgba), vp(lw)) != 0:
            raise RptError(self._err())
        return rgb, rgba, lw

    def bake_importance_map(self, rows: int, cols: int, luminance: np.ndarray, basis: np.ndarray, wavelength_bounds, download: bool = True):
        """ImportanceMap::bake_raw on the device from the scene's resident environment texels (rpt.h N3); installs the tables
        in the scene. -> dict(row_pdf, row_cdf (rows, cols), marginal_pdf, marginal_cdf (rows,), marginal_integral) when
        download, else only marginal_integral."""
        luminance = np.ascontiguousarray(luminance, dtype=F32)
        basis = np.ascontiguousarray(basis, dtype=F32)
        b = RptImapBake()
        b.rows, b.cols, b.num_samples = rows, cols, len(luminance)
        b.lambda_lo, b.lambda_hi = wavelength_bounds
        b.luminance, b.basis = luminance.ctypes.data_as(PF), basis.ctypes.data_as(PF)
        out = {}
        ptrs = [None] * 4
        if download:
            out = dict(row_pdf=np.zeros((rows, cols), dtype=F32), row_cdf=np.zeros((rows, cols), dtype=F32),
                       marginal_pdf=np.zeros(rows, dtype=F32), marginal_cdf=np.zeros(rows, dtype=F32))
            ptrs = [out[k].ctypes.data_as(ct.c_void_p) for k in ("row_pdf", "row_cdf", "marginal_pdf", "marginal_cdf")]
        integral = c_f()
        fn = self._fn("scene_bake_importance_map")
        fn.argtypes = [ct.c_void_p, ct.POINTER(RptImapBake)] + [ct.c_void_p] * 4 + [ct.POINTER(c_f)]
        fn.restype = ct.c_int
        if fn(self.handle, ct.byref(b), *ptrs, ct.byref(integral)) != 0:
            raise RptError(self._err())
        out["marginal_integral"] = float(integral.value)
        return out

    def kernel_times(self) -> List[dict]:
        buf = (RptKernelTime * 32)()
        n = c_u32()
        if self.lib.rpt_last_kernel_times(self.handle, buf, 32, ct.byref(n)) != 0:
            raise RptError(self._err())
        return [{"name": buf[i].name.decode(), "launches": int(buf[i].launches), "ms": float(buf[i].ms)} for i in range(n.value)]

    def stats(self) -> dict:
        s = RptSceneStats()
        if self.lib.rpt_scene_stats(self.handle, ct.byref(s)) != 0:
            raise RptError(self._err())
        return {k: int(getattr(s, k)) for k, _ in s._fields_}

    def close(self) -> None:
        if self.handle:
            self._fn("scene_destroy")(self.handle)
            self.handle = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiScene:
    """RAII wrapper over RptMulti*: one scene replica per device, spp split + film exchange inside the library
    (include/rpt.h rpt_multi_*). This is what `CudaRenderer { devices }` of the Rust shim drives from one host thread."""

    def __init__(self, lib: ct.CDLL, flat: FlatScene, devices):
        self.lib, self.flat, self.devices = lib, flat, list(devices)
        self.handle = ct.c_void_p()
        arr = (ct.c_int * len(self.devices))(*self.devices)
        if lib.rpt_multi_create(ct.byref(flat.desc), arr, len(self.devices), ct.byref(self.handle)) != 0:
            raise RptError(lib.rpt_last_error().decode("utf-8", "replace"))
        self.times = RptMultiTimes()

    def _err(self) -> str:
        return self.lib.rpt_last_error().decode("utf-8", "replace")

    def bake_importance_map(self, rows: int, cols: int, luminance: np.ndarray, basis: np.ndarray, wavelength_bounds) -> None:
        luminance = np.ascontiguousarray(luminance, dtype=F32)
        basis = np.ascontiguousarray(basis, dtype=F32)
        b = RptImapBake()
        b.rows, b.cols, b.num_samples = rows, cols, len(luminance)
        b.lambda_lo, b.lambda_hi = wavelength_bounds
        b.luminance, b.basis = luminance.ctypes.data_as(PF), basis.ctypes.data_as(PF)
        if self.lib.rpt_multi_bake_importance_map(self.handle, ct.byref(b)) != 0:
            raise RptError(self._err())

    def render_pt(self, params: RptRenderParams, film_ptr: Optional[int] = None):
        """-> (film (H, W, 4) f32 mean XYZ or None when film_ptr is given, summed counters). params.spp = TOTAL samples."""
        film = None
        if film_ptr is None:
            film = np.zeros((params.height, params.width, 4), dtype=F32)
            film_ptr = film.ctypes.data
        counters = RptCounters()
        if self.lib.rpt_multi_render_pt(self.handle, ct.byref(params), ct.c_void_p(film_ptr), ct.byref(counters), ct.byref(self.times)) != 0:
            raise RptError(self._err())
        return film, counters

    def close(self) -> None:
        if self.handle:
            self.lib.rpt_multi_destroy(self.handle)
            self.handle = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
