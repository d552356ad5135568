"""`CudaRenderer`: the Python mirror of the Rust shim a maintainer would add to the reference
(`RendererType::Cuda` + `impl Renderer for CudaRenderer`).

Mirrors reference src/renderer/mod.rs:107-112 (`trait Renderer { fn render(&self, world, config) }`)
and the PT branch of src/renderer/naive.rs:410-537 / tiled.rs:553-669: for every render setting whose
integrator is PT it builds the integrator parameters exactly as
`Integrator::from_settings_and_world` does (src/integrator/mod.rs:59-105), then calls
`render_sampled`, which here is one call through the C ABI into the CUDA wavefront pipeline.
There is no CPU fallback: if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import curves as C
from . import ffi
from . import world as W
from .loader import Config, RenderSettings

F32 = np.float32


@dataclass
class PTSettings:
    """PathTracingIntegrator fields (reference src/integrator/pt.rs:16-26)."""

    width: int
    height: int
    min_samples: int
    min_bounces: int
    max_bounces: int
    light_samples: int
    only_direct: bool
    wavelength_bounds: Tuple[float, float]
    camera: int

    @staticmethod
    def from_render_settings(rs: RenderSettings, camera_index: int) -> "PTSettings":
        if rs.integrator_type != "PT":
            raise ValueError("CudaRenderer supports IntegratorKind::PT only (supported_integrators)")
        if rs.medium_aware:
            raise ValueError("medium_aware = true is out of scope (SURVEY.md §8f N1)")
        if rs.max_bounces is None:
            raise ValueError("max_bounces is required (src/integrator/mod.rs:70 unwrap)")
        return PTSettings(
            width=rs.width, height=rs.height, min_samples=rs.min_samples,
            min_bounces=rs.min_bounces if rs.min_bounces is not None else 4,
            max_bounces=rs.max_bounces, light_samples=rs.light_samples,
            only_direct=bool(rs.only_direct) if rs.only_direct is not None else False,
            wavelength_bounds=rs.wavelength_bounds or C.BOUNDED_VISIBLE_RANGE,
            camera=camera_index,
        )

    def to_dict(self) -> dict:
        d = dict(self.__dict__)
        d["wavelength_bounds"] = list(self.wavelength_bounds)
        return d

    @staticmethod
    def from_dict(d: dict) -> "PTSettings":
        d = dict(d)
        d["wavelength_bounds"] = tuple(d["wavelength_bounds"])
        return PTSettings(**d)

    def params(self, seed: int = 0, spp: Optional[int] = None, spp_offset: int = 0, spp_total: Optional[int] = None, flags: int = 0) -> ffi.RptRenderParams:
        p = ffi.RptRenderParams()
        p.width, p.height = self.width, self.height
        p.spp = self.min_samples if spp is None else spp
        p.spp_offset = spp_offset
        p.spp_total = self.min_samples if spp_total is None else spp_total
        p.min_bounces, p.max_bounces = self.min_bounces, self.max_bounces
        p.light_samples, p.only_direct = self.light_samples, int(self.only_direct)
        p.lambda_lo, p.lambda_hi = self.wavelength_bounds
        p.camera = self.camera
        p.seed = seed
        p.flags = flags  # run-time instrumentation (ffi.FLAG_KERNEL_TIMES | ffi.FLAG_BVH_STATS); off by default
        return p


TONEMAPPERS = {"Clamp": 0, "Reinhard0": 1, "Reinhard1": 2}
COLORSPACES = {"sRGB": 0, "Rec709": 1, "Rec2020": 2}


def output_settings(rs: RenderSettings, factor: float = 1.0) -> "ffi.RptOutputSettings":
    """TonemapSettings + ColorSpaceSettings of a render setting (reference src/parsing/tonemap.rs:9-31,
    src/parsing/config.rs:33-43) -> RptOutputSettings; factor is multiplied by `premultiply` as in
    output_film (src/renderer/mod.rs:25)."""
    tm = rs.raw.get("tonemap_settings", {"type": "Clamp", "luminance_only": True})
    cs = rs.raw.get("colorspace_settings", {"type": "sRGB"})
    o = ffi.RptOutputSettings()
    o.tonemapper = TONEMAPPERS[tm["type"]]
    o.luminance_only = int(bool(tm.get("luminance_only", True)))
    o.exposure = float(tm.get("exposure") or 0.0)
    o.key_value = float(tm.get("key_value", 0.18))
    o.white_point = float(tm.get("white_point", 1.0))
    o.colorspace = COLORSPACES[cs["type"]]
    o.factor = float(factor) * float(rs.premultiply if rs.premultiply is not None else 1.0)
    return o


def split_spp(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(count, offset) of the samples-per-pixel share of `rank` (remainder to low ranks)."""
    base, rem = divmod(total, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return count, offset


class CudaRenderer:
    """Drop-in for NaiveRenderer / TiledRenderer on the PT path."""

    def __init__(self, device: int = 0, num_lambda: int = 1024, seed: int = 0):
        self.device, self.num_lambda, self.seed = device, num_lambda, seed
        self.lib = ffi.load_library()
        self._flat_cache: Dict[tuple, tuple] = {}

    def flatten(self, world: W.World, wavelength_bounds: Tuple[float, float]) -> ffi.FlatScene:
        """World -> RptSceneDesc arrays (the shim's flatten.rs). The flattening of an unchanged World is reused: a frame loop
        over the same scene re-uploads it every frame (rpt_scene_create) but does not re-sample its curves on the host."""
        # the key covers what this package itself replaces on a live World (cameras re-aspected per render setting, importance
        # map tables baked / requested); anything else is treated as immutable once flattened
        env = world.environment
        key = (id(world), float(wavelength_bounds[0]), float(wavelength_bounds[1]), self.num_lambda, tuple(id(c) for c in world.cameras),
               id(env.imap_row_pdf), id(env.imap_request), env.kind, len(world.instances), len(world.materials), len(world.curves))
        hit = self._flat_cache.get(key)
        if hit is None or hit[0] is not world:
            if len(self._flat_cache) >= 8:
                self._flat_cache.clear()
            hit = (world, ffi.FlatScene(world, wavelength_bounds[0], wavelength_bounds[1], self.num_lambda))
            self._flat_cache[key] = hit
        return hit[1]

    def supported_integrators(self) -> List[str]:
        return ["PT"]

    def make_scene(self, world: W.World, wavelength_bounds: Tuple[float, float]) -> ffi.Scene:
        flat = self.flatten(world, wavelength_bounds)
        scene = ffi.Scene(self.lib, flat, self.device)
        env = world.environment
        if env.kind == 2 and env.imap_row_pdf is None and env.imap_request is not None:
            # phase 2 of NaiveRenderer::render (src/renderer/naive.rs:469-487): an HDR environment whose importance map is
            # still unbaked is baked before rendering - here on the device, from the texels the scene just uploaded
            from .importance_map import bake_importance_map_on_device

            rows, cols, lum = env.imap_request
            bake_importance_map_on_device(scene, world, rows, cols, lum, wavelength_bounds, download=False)
        return scene

    def make_multi_scene(self, world: W.World, wavelength_bounds: Tuple[float, float], devices) -> "ffi.MultiScene":
        """`RendererType::Cuda { devices }` of the shim: one replica per device + the in-library spp split and film exchange
        (rpt_multi_*); the importance map of an Unbaked HDR environment is baked on every device."""
        flat = self.flatten(world, wavelength_bounds)
        ms = ffi.MultiScene(self.lib, flat, devices)
        env = world.environment
        if env.kind == 2 and env.imap_row_pdf is None and env.imap_request is not None:
            from .importance_map import bake_curve_tables

            rows, cols, lum = env.imap_request
            lum_t, basis_t = bake_curve_tables(world, lum, wavelength_bounds)
            ms.bake_importance_map(rows, cols, lum_t, basis_t, wavelength_bounds)
        return ms

    def render_sampled(self, scene: ffi.Scene, st: PTSettings, spp: Optional[int] = None, spp_offset: int = 0,
                       spp_total: Optional[int] = None):
        """-> (film (H, W, 4) float32 mean XYZ, counters). One C-ABI call; host film out."""
        return scene.render_pt(st.params(self.seed, spp, spp_offset, spp_total))

    def render(self, world: W.World, config: Config, output_dir: Optional[str] = None) -> Dict[str, np.ndarray]:
        """trait Renderer::render: one film per render setting, keyed by filename. With output_dir, each film also goes
        through output_film (tonemap + colour space on the device) and is written as <filename>.png (+ linear .npy)."""
        films = {}
        scenes: Dict[Tuple[float, float], ffi.Scene] = {}
        for i, rs in enumerate(config.render_settings):
            st = PTSettings.from_render_settings(rs, config.camera_names_to_index[rs.camera_id])
            if st.wavelength_bounds not in scenes:
                scenes[st.wavelength_bounds] = self.make_scene(world, st.wavelength_bounds)
            film, _ = self.render_sampled(scenes[st.wavelength_bounds], st)
            films[rs.filename or f"render_{i}"] = film
            if output_dir is not None:
                self.output_film(scenes[st.wavelength_bounds], rs, film, output_dir)
        for s in scenes.values():
            s.close()
        return films

    def output_film(self, scene: ffi.Scene, rs: RenderSettings, film: Optional[np.ndarray], output_dir: Optional[str] = None, factor: float = 1.0):
        """output_film (reference src/renderer/mod.rs:24-80) with the tonemapping / colour conversion / byte encoding on the
        device. film=None tonemaps the device-resident film of the scene's last render. File encoding stays on the host, as
        in the reference: `<filename>.exr` (linear RGB, f32; exr.py) and `<filename>.png` (Pillow; raw .npy bytes without it)."""
        rgb, rgba, lw = scene.output_film(output_settings(rs, factor), film, rs.width, rs.height)
        if output_dir is not None:
            import os

            os.makedirs(output_dir, exist_ok=True)
            name = rs.filename or "beauty"
            from .exr import write_exr_rgb

            write_exr_rgb(os.path.join(output_dir, name + ".exr"), rgb)
            try:
                from PIL import Image

                Image.fromarray(rgba, "RGBA").save(os.path.join(output_dir, name + ".png"))
            except ImportError:
                np.save(os.path.join(output_dir, name + ".rgba8.npy"), rgba)
        return rgb, rgba, lw

    # ---- multi-GPU: spp split + one NCCL reduce of the XYZ film -----------------------
    def render_sampled_distributed(self, scene: ffi.Scene, st: PTSettings, rank: int, world_size: int):
        """Each rank renders its share of min_samples into a device-resident SUM film, then one
        torch.distributed reduce(sum) to rank 0 over NCCL, normalised on the device.
        Returns a torch.cuda tensor (H, W, 4) on every rank (only rank 0 holds the result)."""
        import torch
        import torch.distributed as dist

        count, offset = split_spp(st.min_samples, world_size, rank)
        ptr, counters = scene.render_pt_device(st.params(self.seed, count, offset, 0))
        film = device_tensor(ptr, (st.height, st.width, 4), self.device)
        if world_size > 1:
            dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
            # the collective runs on NCCL's stream and only torch's current stream waits for it; the library normalises on
            # its own stream, so the host has to see the reduce finished first
            torch.cuda.current_stream(self.device).synchronize()
        if rank == 0:
            scene.film_scale(film.data_ptr(), st.height * st.width, 1.0 / st.min_samples)
        return film, counters


def device_tensor(ptr: int, shape, device: int):
    """Zero-copy torch view of a library-owned device buffer (CUDA array interface v3)."""
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}

    return torch.as_tensor(_Arr(), device=f"cuda:{device}")
