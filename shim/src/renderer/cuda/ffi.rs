//! `extern "C"` surface of librpt_b200 (include/rpt.h, RPT_ABI_VERSION 8), field for field.
//!
//! COMPILE-UNVERIFIED: the image this was written in has no Rust toolchain. The layouts are checked instead against the C
//! compiler by tests/test_abi.py (sizeof of every struct vs the ctypes mirror, which has the same field order as this
//! file); a `#[test] fn layout()` at the bottom repeats the sizes so `cargo test` catches drift on the Rust side.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

pub const RPT_ABI_VERSION: u32 = 8;

// MaterialId packing (materials/mod.rs:22-27): (tag << 16) | index
pub const RPT_MAT_TAG_MATERIAL: u32 = 0;
pub const RPT_MAT_TAG_LIGHT: u32 = 1;
pub const RPT_MAT_NONE: u32 = 0xFFFF_FFFF;
pub const fn rpt_mat_pack(tag: u32, idx: u32) -> u32 {
    (tag << 16) | (idx & 0xFFFF)
}

// RptAggregateKind (geometry/mod.rs:17-116)
pub const RPT_AGG_RECT: u32 = 0;
pub const RPT_AGG_SPHERE: u32 = 1;
pub const RPT_AGG_DISK: u32 = 2;
pub const RPT_AGG_MESH: u32 = 3;
// RptAxis (math::Axis)
pub const RPT_AXIS_X: u32 = 0;
pub const RPT_AXIS_Y: u32 = 1;
pub const RPT_AXIS_Z: u32 = 2;
// RptMaterialType
pub const RPT_MATERIAL_LAMBERTIAN: u32 = 0;
pub const RPT_MATERIAL_GGX: u32 = 1;
pub const RPT_MATERIAL_DIFFUSE_LIGHT: u32 = 2;
pub const RPT_MATERIAL_SHARP_LIGHT: u32 = 3;
// RptSidedness (math::Sidedness)
pub const RPT_SIDED_FORWARD: u32 = 0;
pub const RPT_SIDED_REVERSE: u32 = 1;
pub const RPT_SIDED_DUAL: u32 = 2;
// RptEnvKind (world/environment.rs:7-27)
pub const RPT_ENV_CONSTANT: u32 = 0;
pub const RPT_ENV_SUN: u32 = 1;
pub const RPT_ENV_HDR: u32 = 2;
// RptCameraKind
pub const RPT_CAMERA_PROJECTIVE: u32 = 0;
pub const RPT_CAMERA_PANORAMA: u32 = 1;
// RptTonemapper / RptColorSpace (parsing/tonemap.rs:9-31, parsing/config.rs:33-43)
pub const RPT_TONEMAP_CLAMP: u32 = 0;
pub const RPT_TONEMAP_REINHARD0: u32 = 1;
pub const RPT_TONEMAP_REINHARD1: u32 = 2;
pub const RPT_COLORSPACE_SRGB: u32 = 0;
pub const RPT_COLORSPACE_REC709: u32 = 1;
pub const RPT_COLORSPACE_REC2020: u32 = 2;
// RptRenderParams.flags
pub const RPT_FLAG_KERNEL_TIMES: u32 = 1;
pub const RPT_FLAG_BVH_STATS: u32 = 2;
// RptMultiMethod
pub const RPT_MULTI_PEER: u32 = 0;
pub const RPT_MULTI_NCCL: u32 = 1;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RptInstance {
    pub kind: u32,
    pub origin: [f32; 3],
    pub size: [f32; 2],
    pub axis: u32,
    pub two_sided: u32,
    pub mesh: i32,
    pub has_transform: u32,
    pub forward: [f32; 16],
    pub reverse: [f32; 16],
    pub material: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RptMesh {
    pub num_vertices: u32,
    pub num_faces: u32,
    pub vertices: *const f32,
    pub indices: *const u32,
    pub normals: *const f32,
    pub face_material: *const u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RptMaterial {
    pub type_: u32,
    pub texstack: i32,
    pub curve_a: i32,
    pub curve_b: i32,
    pub curve_c: i32,
    pub alpha: f32,
    pub sharpness: f32,
    pub sidedness: u32,
    pub metallic: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RptTexture {
    pub channels: u32,
    pub width: u32,
    pub height: u32,
    pub texels: *const f32,
    pub curves: [i32; 4],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RptTexStack {
    pub first: u32,
    pub count: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RptEnvironment {
    pub kind: u32,
    pub strength: f32,
    pub curve: i32,
    pub angular_diameter: f32,
    pub sun_direction: [f32; 3],
    pub texstack: i32,
    pub rot_forward: [f32; 16],
    pub rot_reverse: [f32; 16],
    pub imap_rows: u32,
    pub imap_cols: u32,
    pub imap_row_pdf: *const f32,
    pub imap_row_cdf: *const f32,
    pub imap_marginal_n: u32,
    pub imap_marginal_pdf: *const f32,
    pub imap_marginal_cdf: *const f32,
    pub imap_marginal_integral: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RptCamera {
    pub origin: [f32; 3],
    pub u: [f32; 3],
    pub v: [f32; 3],
    pub w: [f32; 3],
    pub lower_left: [f32; 3],
    pub horizontal: [f32; 3],
    pub vertical: [f32; 3],
    pub aperture_diameter: f32,
    pub kind: u32,
    pub angle_span: [f32; 2],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RptSceneDesc {
    pub abi_version: u32,
    pub num_instances: u32,
    pub instances: *const RptInstance,
    pub num_meshes: u32,
    pub meshes: *const RptMesh,
    pub num_lights: u32,
    pub lights: *const u32,
    pub num_materials: u32,
    pub materials: *const RptMaterial,
    pub num_curves: u32,
    pub num_lambda: u32,
    pub lut_lambda_lo: f32,
    pub lut_lambda_hi: f32,
    pub curve_lut: *const f32,
    pub cie_lut: *const f32,
    pub num_textures: u32,
    pub textures: *const RptTexture,
    pub num_texstack_textures: u32,
    pub texstack_textures: *const u32,
    pub num_texstacks: u32,
    pub texstacks: *const RptTexStack,
    pub environment: RptEnvironment,
    pub env_sampling_probability: f32,
    pub num_cameras: u32,
    pub cameras: *const RptCamera,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RptRenderParams {
    pub width: u32,
    pub height: u32,
    pub spp: u32,
    pub spp_offset: u32,
    pub spp_total: u32,
    pub min_bounces: u32,
    pub max_bounces: u32,
    pub light_samples: u32,
    pub only_direct: u32,
    pub lambda_lo: f32,
    pub lambda_hi: f32,
    pub camera: u32,
    pub seed: u64,
    pub flags: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RptCounters {
    pub camera_rays: u64,
    pub bounce_rays: u64,
    pub shadow_rays: u64,
    pub light_rays: u64,
    pub env_hits: u64,
    pub segments: u64,
    pub true_rays: u64,
    pub kernel_launches: u64,
    pub shadow_rays_traced: u64,
    pub walk_nodes: u64,
    pub walk_tris: u64,
    pub walk_insts: u64,
    pub shadow_nodes: u64,
    pub shadow_tris: u64,
    pub shadow_insts: u64,
    pub nee_vertices: u64,
    pub device_ms: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RptKernelTime {
    pub name: *const c_char,
    pub launches: u32,
    pub ms: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RptOutputSettings {
    pub tonemapper: u32,
    pub luminance_only: u32,
    pub exposure: f32,
    pub key_value: f32,
    pub white_point: f32,
    pub colorspace: u32,
    pub factor: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RptImapBake {
    pub rows: u32,
    pub cols: u32,
    pub num_samples: u32,
    pub lambda_lo: f32,
    pub lambda_hi: f32,
    pub luminance: *const f32,
    pub basis: *const f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RptSceneStats {
    pub tlas_nodes: u64,
    pub blas_nodes: u64,
    pub triangles: u64,
    pub instances: u64,
    pub node_bytes: u64,
    pub triangle_bytes: u64,
    pub scene_bytes_total: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RptMultiTimes {
    pub method: u32,
    pub devices: u32,
    pub render_device_ms_max: f64,
    pub exchange_device_ms: f64,
    pub render_wall_ms: f64,
    pub exchange_wall_ms: f64,
    pub download_wall_ms: f64,
}

#[repr(C)]
pub struct RptScene {
    _private: [u8; 0],
}
#[repr(C)]
pub struct RptMulti {
    _private: [u8; 0],
}

#[link(name = "rpt_b200")]
extern "C" {
    pub fn rpt_last_error() -> *const c_char;
    pub fn rpt_abi_version() -> u32;
    pub fn rpt_device_count(count: *mut c_int) -> c_int;
    pub fn rpt_scene_create(desc: *const RptSceneDesc, device: c_int, out: *mut *mut RptScene) -> c_int;
    pub fn rpt_scene_destroy(scene: *mut RptScene) -> c_int;
    pub fn rpt_render_pt(scene: *mut RptScene, params: *const RptRenderParams, film_xyzw: *mut f32, counters: *mut RptCounters) -> c_int;
    pub fn rpt_render_pt_device(scene: *mut RptScene, params: *const RptRenderParams, film_dev: *mut *mut c_void, counters: *mut RptCounters) -> c_int;
    pub fn rpt_trace_primary(scene: *mut RptScene, params: *const RptRenderParams, instance_id: *mut u32, primitive_id: *mut u32, t: *mut f32) -> c_int;
    pub fn rpt_trace_rays(scene: *mut RptScene, n: u32, origins: *const f32, dirs: *const f32, tmax: *const f32, instance_id: *mut u32,
                          primitive_id: *mut u32, t: *mut f32) -> c_int;
    pub fn rpt_film_scale(scene: *mut RptScene, film_dev: *mut c_void, n_float4: u64, scale: f32) -> c_int;
    pub fn rpt_last_kernel_times(scene: *mut RptScene, out: *mut RptKernelTime, cap: u32, n: *mut u32) -> c_int;
    pub fn rpt_output_film(scene: *mut RptScene, film_xyzw: *const f32, width: u32, height: u32, settings: *const RptOutputSettings,
                           rgb_linear: *mut f32, rgba8: *mut u8, l_w: *mut f32) -> c_int;
    pub fn rpt_scene_bake_importance_map(scene: *mut RptScene, bake: *const RptImapBake, row_pdf: *mut f32, row_cdf: *mut f32,
                                         marginal_pdf: *mut f32, marginal_cdf: *mut f32, marginal_integral: *mut f32) -> c_int;
    pub fn rpt_multi_create(desc: *const RptSceneDesc, devices: *const c_int, n: c_int, out: *mut *mut RptMulti) -> c_int;
    pub fn rpt_multi_destroy(multi: *mut RptMulti) -> c_int;
    pub fn rpt_multi_scene(multi: *mut RptMulti, index: c_int, scene: *mut *mut RptScene) -> c_int;
    pub fn rpt_multi_bake_importance_map(multi: *mut RptMulti, bake: *const RptImapBake) -> c_int;
    pub fn rpt_multi_render_pt(multi: *mut RptMulti, params: *const RptRenderParams, film_xyzw: *mut f32, counters: *mut RptCounters,
                               times: *mut RptMultiTimes) -> c_int;
    pub fn rpt_render_pt_multi(desc: *const RptSceneDesc, devices: *const c_int, n: c_int, params: *const RptRenderParams, film_xyzw: *mut f32,
                               counters: *mut RptCounters) -> c_int;
    pub fn rpt_probe_bandwidth(device: c_int, bytes: u64, reps: u32, mode: c_int, gbps: *mut f64) -> c_int;
    pub fn rpt_scene_stats(scene: *mut RptScene, out: *mut RptSceneStats) -> c_int;
}

/// `rpt_last_error()` as an owned String.
pub fn last_error() -> String {
    unsafe {
        let p = rpt_last_error();
        if p.is_null() {
            String::new()
        } else {
            std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use std::mem::size_of;

    /// sizes printed by `gcc -I include` for include/rpt.h on x86-64 (tests/test_abi.py::test_struct_sizes_match_the_c_layout)
    #[test]
    fn layout() {
        assert_eq!(size_of::<RptInstance>(), 172);
        assert_eq!(size_of::<RptMesh>(), 40);
        assert_eq!(size_of::<RptMaterial>(), 36);
        assert_eq!(size_of::<RptTexture>(), 40);
        assert_eq!(size_of::<RptEnvironment>(), 216);
        assert_eq!(size_of::<RptCamera>(), 100);
        assert_eq!(size_of::<RptSceneDesc>(), 376);
        assert_eq!(size_of::<RptRenderParams>(), 64);
        assert_eq!(size_of::<RptCounters>(), 136);
        assert_eq!(size_of::<RptSceneStats>(), 56);
        assert_eq!(size_of::<RptOutputSettings>(), 28);
        assert_eq!(size_of::<RptImapBake>(), 40);
        assert_eq!(size_of::<RptKernelTime>(), 16);
        assert_eq!(size_of::<RptMultiTimes>(), 48);
    }
}
