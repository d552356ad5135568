"""rust-pathtracer_b200 — B200-native backend for the PT integrator path of
gillett-hernandez/rust-pathtracer (reference src/integrator/pt.rs and its per-ray callees).

The directory name carries a hyphen (it mirrors the reference's repository name), so it is
imported through `__graft_entry__.load_package()` under the module name `rust_pathtracer_b200`.

Layout: csrc/ (CUDA kernels + the C ABI of include/rpt.h), ffi.py (ctypes binding + World
flattening), exr.py (EXR file encoding of output_film's payload), loader.py / curves.py / world.py / importance_map.py (host mirror of the reference's
parsing + World), renderer.py (`CudaRenderer`, the mirror of the reference's `Renderer` trait),
blob.py (portable flattened scenes).
"""
from . import blob, curves, exr, ffi, importance_map, loader, renderer, world  # noqa: F401
from .renderer import CudaRenderer, PTSettings, split_spp  # noqa: F401

__all__ = ["blob", "curves", "exr", "ffi", "importance_map", "loader", "renderer", "world", "CudaRenderer", "PTSettings", "split_spp"]
