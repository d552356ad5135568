//! `CudaRenderer`: the B200 backend behind the reference's own `Renderer` plug point
//! (`trait Renderer { fn render(&self, world: World, config: &Config) }`, src/renderer/mod.rs:107-112), selected by
//! `RendererType::Cuda { devices }` in `[renderer]` (src/parsing/config.rs:109-121) through `construct_renderer`
//! (src/bin/main.rs:59-68). It repeats phases 1-2 of `NaiveRenderer::render` (src/renderer/naive.rs:410-537) for the
//! PathTracing settings and replaces `render_sampled` (naive.rs:27-119 / tiled.rs:279-542) by one call into librpt_b200.
//!
//! COMPILE-UNVERIFIED: written in an image without a Rust toolchain. Its behaviour is what
//! `rust-pathtracer_b200/renderer.py::CudaRenderer` does (the Python mirror every GPU test drives).
mod ffi;
mod flatten;

use std::os::raw::c_int;

use crate::parsing::config::{Config, IntegratorKind, RenderSettings};
use crate::prelude::*;
use crate::renderer::{output_film, Renderer};
use crate::world::World;
use math::spectral::BOUNDED_VISIBLE_RANGE;

use self::ffi::*;
use self::flatten::FlatWorld;

pub struct CudaRenderer {
    /// CUDA ordinals; empty = every visible device. devices[0] receives the reduced film.
    pub devices: Vec<u32>,
    /// Philox key (the reference's RandomSampler is OS-seeded and not reproducible, naive.rs:79; this backend is)
    pub seed: u64,
}

impl CudaRenderer {
    pub fn new(devices: Option<Vec<u32>>) -> Self {
        CudaRenderer { devices: devices.unwrap_or_default(), seed: 0 }
    }

    fn device_list(&self) -> Vec<c_int> {
        if !self.devices.is_empty() {
            return self.devices.iter().map(|&d| d as c_int).collect();
        }
        let mut n: c_int = 0;
        if unsafe { rpt_device_count(&mut n) } != 0 || n < 1 {
            panic!("rpt_device_count: {} (there is no CPU fallback)", last_error());
        }
        (0..n).collect()
    }

    /// `render_sampled` for one PathTracing render setting: flatten, upload to every device, bake an Unbaked importance map on
    /// the devices, split `min_samples` over the devices, one film exchange over NVLink, mean XYZ film back.
    fn render_sampled(&self, world: &World, cameras: &[CameraEnum], camera_index: usize, settings: &RenderSettings) -> Vec2D<XYZColor> {
        let (light_samples, medium_aware) = match settings.integrator {
            IntegratorKind::PT { light_samples, medium_aware } => (light_samples, medium_aware),
            _ => unreachable!("supported_integrators() lists PT only"),
        };
        if medium_aware {
            // random_walk_medium cannot finish a path that reaches a lit light (pt.rs:575-581 panics): nothing to be identical to
            panic!("CudaRenderer: medium_aware = true is not supported");
        }
        let bounds = settings.wavelength_bounds.map(|e| Bounds1D::new(e.0, e.1)).unwrap_or(BOUNDED_VISIBLE_RANGE); // integrator/mod.rs:65-68
        let flat = FlatWorld::new(world, cameras, bounds).unwrap_or_else(|e| panic!("CudaRenderer: {}", e));
        let desc = flat.desc();
        let devices = self.device_list();
        let mut multi: *mut RptMulti = std::ptr::null_mut();
        if unsafe { rpt_multi_create(&desc, devices.as_ptr(), devices.len() as c_int, &mut multi) } != 0 {
            panic!("rpt_multi_create: {}", last_error());
        }
        // phase 2 of NaiveRenderer::render (naive.rs:469-487): bake the importance map if the environment still carries an
        // Unbaked one - here on the devices, from the texels rpt_multi_create just uploaded
        if let Some((rows, cols, luminance)) = &flat.unbaked_importance_map {
            if world.get_env_sampling_probability() > 0.0 {
                let (lum, basis) = flat.imap_bake_tables(luminance, 100);
                let bake = RptImapBake { rows: *rows as u32, cols: *cols as u32, num_samples: 100, lambda_lo: bounds.lower, lambda_hi: bounds.upper,
                                         luminance: lum.as_ptr(), basis: basis.as_ptr() };
                if unsafe { rpt_multi_bake_importance_map(multi, &bake) } != 0 {
                    panic!("rpt_multi_bake_importance_map: {}", last_error());
                }
            }
        }
        let (width, height) = (settings.resolution.width, settings.resolution.height);
        let params = RptRenderParams {
            width: width as u32,
            height: height as u32,
            spp: settings.min_samples as u32,       // the library splits this over the devices
            spp_offset: 0,
            spp_total: settings.min_samples as u32, // mean over min_samples (tiled.rs:396-398)
            min_bounces: settings.min_bounces.unwrap_or(4) as u32, // integrator/mod.rs:96
            max_bounces: settings.max_bounces.unwrap() as u32,     // integrator/mod.rs:70
            light_samples: light_samples as u32,
            only_direct: settings.only_direct.unwrap_or(false) as u32,
            lambda_lo: bounds.lower,
            lambda_hi: bounds.upper,
            camera: camera_index as u32,
            seed: self.seed,
            flags: 0,
            reserved: 0,
        };
        let mut film = vec![0.0f32; width * height * 4];
        let mut counters = RptCounters::default();
        let mut times = RptMultiTimes::default();
        let rc = unsafe { rpt_multi_render_pt(multi, &params, film.as_mut_ptr(), &mut counters, &mut times) };
        let err = if rc != 0 { Some(last_error()) } else { None };
        unsafe { rpt_multi_destroy(multi) };
        if let Some(e) = err {
            panic!("rpt_multi_render_pt: {}", e);
        }
        // Profile::pretty_print's numbers (src/profile.rs:36-80), from the device counters
        let total = counters.camera_rays + counters.bounce_rays + counters.shadow_rays + counters.light_rays;
        info!(
            "took {:.3} ms on {} device(s) ({} exchange {:.3} ms): {} camera, {} bounce, {} shadow rays; {:.1} Mrays/s (reference definition), {:.1} M segments/s",
            counters.device_ms, times.devices, if times.method == RPT_MULTI_NCCL { "NCCL" } else { "NVLink peer" }, times.exchange_device_ms,
            counters.camera_rays, counters.bounce_rays, counters.shadow_rays,
            total as f64 / counters.device_ms / 1e3, counters.segments as f64 / counters.device_ms / 1e3
        );
        let mut out = Vec2D::new(width, height, XYZColor::BLACK);
        for (px, c) in out.buffer.iter_mut().zip(film.chunks_exact(4)) {
            *px = XYZColor::new(c[0], c[1], c[2]);
        }
        out
    }
}

impl Renderer for CudaRenderer {
    fn render(&self, world: World, config: &Config) {
        // phase 1 (naive.rs:424-463): one aspect-corrected camera per render setting; only PathTracing settings are taken
        let mut cameras: Vec<CameraEnum> = Vec::new();
        let mut jobs: Vec<(usize, RenderSettings)> = Vec::new();
        for settings in config.render_settings.iter() {
            if !matches!(settings.integrator, IntegratorKind::PT { .. }) {
                warn!("CudaRenderer skips render setting {:?}: only IntegratorKind::PT is supported", settings.filename);
                continue;
            }
            let aspect_ratio = settings.resolution.width as f32 / settings.resolution.height as f32;
            let camera = world.cameras[config.camera_names_to_index[&settings.camera_id]].clone().with_aspect_ratio(aspect_ratio);
            cameras.push(camera);
            jobs.push((cameras.len() - 1, settings.clone()));
        }
        // phase 2 (naive.rs:466-509): render each, output as soon as it is finished
        for (camera_index, settings) in jobs.iter() {
            let film = self.render_sampled(&world, &cameras, *camera_index, settings);
            output_film(settings, &film, 1.0); // tonemap + EXR/PNG exactly as the CPU renderers do (renderer/mod.rs:24-80)
        }
        // (rpt_output_film can produce the same PNG / EXR payloads on the device from the film the root still holds; a
        //  maintainer who wants to skip the 133 MB 4K film download calls it through rpt_multi_scene(multi, 0) instead.)
    }
    fn supported_integrators(&self) -> &[IntegratorKind] {
        &[IntegratorKind::PT { light_samples: 0, medium_aware: false }]
    }
}
