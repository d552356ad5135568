//! `World` -> `RptSceneDesc`: the flattening the C ABI asks for (include/rpt.h "Scene: the flattened World").
//!
//! COMPILE-UNVERIFIED (no Rust toolchain in the image this was written in). The executable specification of this file is
//! `rust-pathtracer_b200/ffi.py::FlatScene` (the same flattening from the Python mirror of `World`), which every GPU parity
//! test goes through; the two are kept field for field.
//!
//! Reference types read here: `World` (src/world/mod.rs:18-28), `Accelerator` (src/accelerator/mod.rs:20-28), `Instance`
//! (src/geometry/instance.rs:9-15), `Aggregate::{AARect, Sphere, Disk, Mesh}` (src/geometry/*.rs), `MaterialEnum`
//! (src/materials/mod.rs:141-283), `TexStack / Texture1 / Texture4` (src/texture.rs), `EnvironmentMap`
//! (src/world/environment.rs:7-27), `ImportanceMap` (src/world/importance_map.rs:32-46), `CameraEnum` (src/camera/mod.rs).
//!
//! Every `Curve` / `CurveWithCDF` the path evaluates is sampled HERE with the real `evaluate_power` on a uniform grid over
//! the render's wavelength bounds, so the device never needs `Curve` semantics (SURVEY §0 consequence ii).
//!
//! Visibility: the struct fields marked `pub(crate)?` below are private in the reference today and need `pub(crate)` (or a
//! getter) for this module to compile: `ProjectiveCamera::{u, v, w, lower_left_corner, horizontal, vertical}`
//! (src/camera/projective_camera.rs:16-23) and `PanoramaCamera::transform` (src/camera/panorama_camera.rs:15).
use std::collections::HashMap;
use std::sync::Arc;

use crate::prelude::*;
use crate::texture::{TexStack, Texture};
use crate::world::{Accelerator, Aggregate, EnvironmentMap, ImportanceMap, Instance, MaterialEnum, MaterialId, World};
use math::curves::{Curve, CurveWithCDF};
use math::spectral::x_bar;
use math::spectral::y_bar;
use math::spectral::z_bar;

use super::ffi::*;

/// Number of wavelength samples of every curve LUT (the Python mirror and every committed measurement use 1024).
pub const NUM_LAMBDA: usize = 1024;

/// Owns every array an `RptSceneDesc` points into. Keep it alive across `rpt_scene_create` / `rpt_multi_create`
/// (the library copies what it needs during that call).
pub struct FlatWorld {
    instances: Vec<RptInstance>,
    mesh_vertices: Vec<Vec<f32>>,
    mesh_indices: Vec<Vec<u32>>,
    mesh_normals: Vec<Vec<f32>>,
    mesh_materials: Vec<Vec<u32>>,
    meshes: Vec<RptMesh>,
    lights: Vec<u32>,
    materials: Vec<RptMaterial>,
    curve_lut: Vec<f32>,
    num_curves: usize,
    cie_lut: Vec<f32>,
    texels: Vec<Vec<f32>>,
    textures: Vec<RptTexture>,
    texstack_textures: Vec<u32>,
    texstacks: Vec<RptTexStack>,
    imap_row_pdf: Vec<f32>,
    imap_row_cdf: Vec<f32>,
    imap_marginal_pdf: Vec<f32>,
    imap_marginal_cdf: Vec<f32>,
    cameras: Vec<RptCamera>,
    environment: RptEnvironment,
    env_sampling_probability: f32,
    bounds: Bounds1D,
    /// `Some((rows, cols, luminance_curve))` when the environment's importance map is still `Unbaked`: the caller bakes it on
    /// the device (`rpt_scene_bake_importance_map`) right after scene creation, as naive.rs:469-487 bakes it on the host.
    pub unbaked_importance_map: Option<(usize, usize, Curve)>,
    /// curves of the environment's texture stack, in stack order, 4 per texture (Texture1 uses slot 0): what RptImapBake::basis is evaluated from
    pub env_basis_curves: Vec<[Option<Curve>; 4]>,
}

fn pack_material(id: MaterialId) -> u32 {
    match id {
        MaterialId::Material(i) => rpt_mat_pack(RPT_MAT_TAG_MATERIAL, i as u32),
        MaterialId::Light(i) => rpt_mat_pack(RPT_MAT_TAG_LIGHT, i as u32),
        // MaterialId::Camera only ever sits on camera lens / surface instances, which are not part of World.accelerator
        MaterialId::Camera(i) => rpt_mat_pack(2, i as u32),
    }
}

/// Row-major 4x4 of a `Matrix4x4`, recovered through the operations the reference itself uses on it (`M * Vec3`,
/// `M * Point3`: src/aabb.rs:116-138, src/geometry/instance.rs:108): column k = M * e_k, translation = M * origin.
fn mat16(m: &Matrix4x4) -> [f32; 16] {
    let cx = *m * Vec3::X;
    let cy = *m * Vec3::Y;
    let cz = *m * Vec3::Z;
    let t = *m * Point3::ORIGIN;
    [
        cx.x(), cy.x(), cz.x(), t.x(),
        cx.y(), cy.y(), cz.y(), t.y(),
        cx.z(), cy.z(), cz.z(), t.z(),
        0.0, 0.0, 0.0, 1.0,
    ]
}

const IDENTITY16: [f32; 16] = [1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0];

fn axis_id(a: &Axis) -> u32 {
    match a {
        Axis::X => RPT_AXIS_X,
        Axis::Y => RPT_AXIS_Y,
        Axis::Z => RPT_AXIS_Z,
    }
}

fn sidedness_id(s: &math::Sidedness) -> u32 {
    match s {
        math::Sidedness::Forward => RPT_SIDED_FORWARD,
        math::Sidedness::Reverse => RPT_SIDED_REVERSE,
        math::Sidedness::Dual => RPT_SIDED_DUAL,
    }
}

/// LUT grid point i of n over [lo, hi] (include/rpt.h `curve_lut`; ffi.py `lut_grid`): f32 arithmetic, inclusive ends.
fn grid(bounds: Bounds1D, n: usize, i: usize) -> f32 {
    bounds.lower + (bounds.upper - bounds.lower) * (i as f32 / (n - 1) as f32)
}

struct CurveTable {
    lut: Vec<f32>,
    count: usize,
    bounds: Bounds1D,
}

impl CurveTable {
    /// Samples `curve` with the real `evaluate_power` and returns its LUT id.
    fn add(&mut self, curve: &Curve) -> i32 {
        for i in 0..NUM_LAMBDA {
            self.lut.push(curve.evaluate_power(grid(self.bounds, NUM_LAMBDA, i)));
        }
        self.count += 1;
        (self.count - 1) as i32
    }
    fn add_cdf(&mut self, curve: &CurveWithCDF) -> i32 {
        // CurveWithCDF::evaluate_power evaluates its pdf curve (texture.rs:104-109 uses curves[i].evaluate_power)
        for i in 0..NUM_LAMBDA {
            self.lut.push(curve.evaluate_power(grid(self.bounds, NUM_LAMBDA, i)));
        }
        self.count += 1;
        (self.count - 1) as i32
    }
}

impl FlatWorld {
    /// Flattens `world` for renders over `bounds` (RenderSettings::wavelength_bounds or BOUNDED_VISIBLE_RANGE,
    /// src/integrator/mod.rs:65-68). `cameras` = the aspect-corrected cameras, one per render setting, in the order the
    /// caller will index them (RptRenderParams::camera).
    pub fn new(world: &World, cameras: &[CameraEnum], bounds: Bounds1D) -> Result<FlatWorld, String> {
        let mut curves = CurveTable { lut: Vec::new(), count: 0, bounds };

        // ---- textures: every TexStack that a material or the environment refers to gets a stack id
        let mut texels: Vec<Vec<f32>> = Vec::new();
        let mut textures: Vec<RptTexture> = Vec::new();
        let mut texstack_textures: Vec<u32> = Vec::new();
        let mut texstacks: Vec<RptTexStack> = Vec::new();
        let mut add_stack = |stack: &TexStack, curves: &mut CurveTable| -> i32 {
            let first = texstack_textures.len() as u32;
            for tex in stack.textures.iter() {
                match tex {
                    Texture::Texture1(t) => {
                        // Vec2D<f32>, row-major y * width + x (src/vec2d.rs:31-32)
                        texels.push(t.texture.buffer.clone());
                        let c = curves.add_cdf(&t.curve);
                        textures.push(RptTexture { channels: 1, width: t.texture.width as u32, height: t.texture.height as u32,
                                                   texels: std::ptr::null(), curves: [c, -1, -1, -1] });
                    }
                    Texture::Texture4(t) => {
                        let mut flat = Vec::with_capacity(t.texture.buffer.len() * 4);
                        for px in t.texture.buffer.iter() {
                            flat.extend_from_slice(&px.to_array());
                        }
                        texels.push(flat);
                        let c = [curves.add_cdf(&t.curves[0]), curves.add_cdf(&t.curves[1]), curves.add_cdf(&t.curves[2]), curves.add_cdf(&t.curves[3])];
                        textures.push(RptTexture { channels: 4, width: t.texture.width as u32, height: t.texture.height as u32,
                                                   texels: std::ptr::null(), curves: c });
                    }
                }
                texstack_textures.push((textures.len() - 1) as u32);
            }
            texstacks.push(RptTexStack { first, count: stack.textures.len() as u32 });
            (texstacks.len() - 1) as i32
        };

        // ---- materials: table order preserved, index 0 is the mauve error light (src/parsing/mod.rs:440-444)
        let mut materials: Vec<RptMaterial> = Vec::with_capacity(world.materials.len());
        for m in world.materials.iter() {
            let mut r = RptMaterial { texstack: -1, curve_a: -1, curve_b: -1, curve_c: -1, ..Default::default() };
            match m {
                MaterialEnum::Lambertian(l) => {
                    r.type_ = RPT_MATERIAL_LAMBERTIAN;
                    r.texstack = add_stack(&l.texture, &mut curves);
                }
                MaterialEnum::GGX(g) => {
                    r.type_ = RPT_MATERIAL_GGX;
                    r.alpha = g.alpha;
                    r.curve_a = curves.add(&g.eta);
                    r.curve_b = curves.add(&g.eta_o);
                    r.curve_c = curves.add(&g.kappa);
                    // GGX::metallic is private; this is the expression that sets it (src/materials/ggx.rs:205)
                    r.metallic = (g.kappa.evaluate_integral(math::spectral::BOUNDED_VISIBLE_RANGE, 100, false) > 0.0) as u32;
                }
                MaterialEnum::DiffuseLight(d) => {
                    r.type_ = RPT_MATERIAL_DIFFUSE_LIGHT;
                    r.curve_a = curves.add(&d.bounce_color);
                    r.curve_b = curves.add_cdf(&d.emit_color);
                    r.sidedness = sidedness_id(&d.sidedness);
                }
                MaterialEnum::SharpLight(s) => {
                    r.type_ = RPT_MATERIAL_SHARP_LIGHT;
                    r.curve_a = curves.add(&s.bounce_color);
                    r.curve_b = curves.add_cdf(&s.emit_color);
                    r.sidedness = sidedness_id(&s.sidedness);
                    r.sharpness = s.sharpness; // already 1 + |sharpness| (src/materials/sharp_light.rs:25)
                }
                #[allow(unreachable_patterns)]
                _ => return Err("material kind outside the PT path's scope (PassthroughFilter is not compiled, src/materials/mod.rs:11)".into()),
            }
            materials.push(r);
        }

        // ---- instances (array index == instance_id) and meshes (deduplicated: the parser clones one Mesh per instance,
        // src/parsing/mod.rs:530-535, sharing the Arc'd vertex / index arrays)
        let inst_list: &Vec<Instance> = match &world.accelerator {
            Accelerator::List { instances } => instances,
            Accelerator::BVH { instances, .. } => instances,
        };
        let n_inst = inst_list.len();
        let mut instances: Vec<Option<RptInstance>> = vec![None; n_inst];
        let mut mesh_ids: HashMap<*const Vec<Point3>, i32> = HashMap::new();
        let (mut mesh_vertices, mut mesh_indices, mut mesh_normals, mut mesh_materials) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for inst in inst_list.iter() {
            let id = inst.instance_id as usize;
            if id >= n_inst || instances[id].is_some() {
                return Err(format!("instance ids are not a permutation of 0..{} (id {})", n_inst, id));
            }
            let mut r = RptInstance { kind: 0, origin: [0.0; 3], size: [0.0; 2], axis: RPT_AXIS_Z, two_sided: 0, mesh: -1, has_transform: 0,
                                      forward: IDENTITY16, reverse: IDENTITY16, material: inst.material_id.map(pack_material).unwrap_or(RPT_MAT_NONE) };
            if let Some(t) = &inst.transform {
                r.has_transform = 1;
                r.forward = mat16(&t.forward);
                r.reverse = mat16(&t.reverse);
            }
            match &inst.aggregate {
                Aggregate::AARect(a) => {
                    r.kind = RPT_AGG_RECT;
                    r.origin = [a.origin.x(), a.origin.y(), a.origin.z()];
                    r.size = [a.size.0, a.size.1];
                    r.axis = axis_id(&a.normal);
                    r.two_sided = a.two_sided as u32;
                }
                Aggregate::Sphere(s) => {
                    r.kind = RPT_AGG_SPHERE;
                    r.origin = [s.origin.x(), s.origin.y(), s.origin.z()];
                    r.size = [s.radius, 0.0];
                }
                Aggregate::Disk(d) => {
                    r.kind = RPT_AGG_DISK;
                    r.origin = [d.origin.x(), d.origin.y(), d.origin.z()];
                    r.size = [d.radius, 0.0];
                    r.two_sided = d.two_sided as u32;
                }
                Aggregate::Mesh(m) => {
                    r.kind = RPT_AGG_MESH;
                    let key = Arc::as_ptr(&m.vertices);
                    let next = mesh_ids.len() as i32;
                    let mid = *mesh_ids.entry(key).or_insert(next);
                    if mid == next {
                        mesh_vertices.push(m.vertices.iter().flat_map(|p| [p.x(), p.y(), p.z()]).collect::<Vec<f32>>());
                        mesh_indices.push(m.indices.iter().map(|&i| i as u32).collect::<Vec<u32>>());
                        // Mesh::normals is empty when the OBJ has none (src/geometry/mesh.rs:169-179 checks len)
                        mesh_normals.push(m.normals.iter().flat_map(|n| [n.x(), n.y(), n.z()]).collect::<Vec<f32>>());
                        mesh_materials.push(m.material_ids.iter().map(|&id| pack_material(id)).collect::<Vec<u32>>());
                    }
                    r.mesh = mid;
                }
            }
            instances[id] = Some(r);
        }
        let instances: Vec<RptInstance> = instances.into_iter().map(|i| i.unwrap()).collect();
        let mut meshes = Vec::with_capacity(mesh_vertices.len());
        for k in 0..mesh_vertices.len() {
            let has_n = !mesh_normals[k].is_empty();
            let has_m = !mesh_materials[k].is_empty();
            meshes.push(RptMesh {
                num_vertices: (mesh_vertices[k].len() / 3) as u32,
                num_faces: (mesh_indices[k].len() / 3) as u32,
                vertices: mesh_vertices[k].as_ptr(),
                indices: mesh_indices[k].as_ptr(),
                normals: if has_n { mesh_normals[k].as_ptr() } else { std::ptr::null() },
                face_material: if has_m { mesh_materials[k].as_ptr() } else { std::ptr::null() }, // empty => Material(0) (mesh.rs:71-75)
            });
        }

        // ---- environment
        let mut env = RptEnvironment {
            kind: RPT_ENV_CONSTANT, strength: 0.0, curve: -1, angular_diameter: 0.0, sun_direction: [0.0, 0.0, 1.0], texstack: -1,
            rot_forward: IDENTITY16, rot_reverse: IDENTITY16, imap_rows: 0, imap_cols: 0, imap_row_pdf: std::ptr::null(),
            imap_row_cdf: std::ptr::null(), imap_marginal_n: 0, imap_marginal_pdf: std::ptr::null(), imap_marginal_cdf: std::ptr::null(),
            imap_marginal_integral: 1.0,
        };
        let (mut row_pdf, mut row_cdf, mut m_pdf, mut m_cdf) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        let mut unbaked = None;
        let mut env_basis_curves = Vec::new();
        match &world.environment {
            EnvironmentMap::Constant { color, strength } => {
                env.kind = RPT_ENV_CONSTANT;
                env.strength = *strength;
                env.curve = curves.add_cdf(color);
            }
            EnvironmentMap::Sun { color, strength, angular_diameter, sun_direction } => {
                env.kind = RPT_ENV_SUN;
                env.strength = *strength;
                env.curve = curves.add_cdf(color);
                env.angular_diameter = *angular_diameter;
                env.sun_direction = [sun_direction.x(), sun_direction.y(), sun_direction.z()];
            }
            EnvironmentMap::HDR { texture, importance_map, rotation, strength } => {
                env.kind = RPT_ENV_HDR;
                env.strength = *strength;
                env.texstack = add_stack(texture, &mut curves);
                env.rot_forward = mat16(&rotation.forward);
                env.rot_reverse = mat16(&rotation.reverse);
                for tex in texture.textures.iter() {
                    env_basis_curves.push(match tex {
                        Texture::Texture1(t) => [Some(t.curve.pdf.clone()), None, None, None],
                        Texture::Texture4(t) => [Some(t.curves[0].pdf.clone()), Some(t.curves[1].pdf.clone()), Some(t.curves[2].pdf.clone()), Some(t.curves[3].pdf.clone())],
                    });
                }
                match importance_map {
                    ImportanceMap::Baked { vertical_resolution, horizontal_resolution, data, marginal_cdf, .. } => {
                        // every row: CurveWithCDF { pdf: Linear Nearest over (0,1), cdf: Linear Nearest } (importance_map.rs:158-176)
                        env.imap_rows = *vertical_resolution as u32;
                        env.imap_cols = *horizontal_resolution as u32;
                        for row in data.iter() {
                            row_pdf.extend_from_slice(linear_signal(&row.pdf)?);
                            row_cdf.extend_from_slice(linear_signal(&row.cdf)?);
                        }
                        m_pdf.extend_from_slice(linear_signal(&marginal_cdf.pdf)?);
                        m_cdf.extend_from_slice(linear_signal(&marginal_cdf.cdf)?);
                        env.imap_marginal_n = m_cdf.len() as u32;
                        env.imap_marginal_integral = marginal_cdf.pdf_integral;
                    }
                    ImportanceMap::Unbaked { luminance_curve, vertical_resolution, horizontal_resolution } => {
                        unbaked = Some((*vertical_resolution, *horizontal_resolution, luminance_curve.clone()));
                    }
                    ImportanceMap::Empty => {}
                }
            }
        }

        // ---- cameras
        let mut cams = Vec::with_capacity(cameras.len());
        for c in cameras.iter() {
            let v3 = |v: Vec3| [v.x(), v.y(), v.z()];
            let p3 = |p: Point3| [p.x(), p.y(), p.z()];
            cams.push(match c {
                CameraEnum::ProjectiveCamera(p) => RptCamera {
                    origin: p3(p.origin), u: v3(p.u), v: v3(p.v), w: v3(p.w), // pub(crate)?
                    lower_left: p3(p.lower_left_corner), horizontal: v3(p.horizontal), vertical: v3(p.vertical), // pub(crate)?
                    aperture_diameter: p.aperture_diameter, kind: RPT_CAMERA_PROJECTIVE, angle_span: [0.0, 0.0],
                },
                CameraEnum::PanoramaCamera(p) => {
                    // PanoramaCamera::get_ray maps the local direction through `transform` (panorama_camera.rs:68-91): the
                    // camera frame is its image of the axes; w = +direction (panorama_camera.rs:31-33)
                    let t = &p.transform; // pub(crate)?
                    RptCamera {
                        origin: p3(p.origin), u: v3(t.to_world(Vec3::X)), v: v3(t.to_world(Vec3::Y)), w: v3(t.to_world(Vec3::Z)),
                        lower_left: [0.0; 3], horizontal: [0.0; 3], vertical: [0.0; 3], aperture_diameter: 0.0,
                        kind: RPT_CAMERA_PANORAMA, angle_span: [p.angle_span.0, p.angle_span.1],
                    }
                }
                #[allow(unreachable_patterns)]
                _ => return Err("RealisticCamera is outside the PT path's scope (cargo feature `realistic_camera`, rust_optics lens tracer)".into()),
            });
        }

        // ---- CIE colour-matching LUT on the same grid (math::spectral::{x,y,z}_bar take angstroms: importance_map.rs:364)
        let mut cie = vec![0.0f32; 3 * NUM_LAMBDA];
        for i in 0..NUM_LAMBDA {
            let a = grid(bounds, NUM_LAMBDA, i) * 10.0;
            cie[i] = x_bar(a);
            cie[NUM_LAMBDA + i] = y_bar(a);
            cie[2 * NUM_LAMBDA + i] = z_bar(a);
        }

        let mut flat = FlatWorld {
            instances, mesh_vertices, mesh_indices, mesh_normals, mesh_materials, meshes,
            lights: world.lights.iter().map(|&l| l as u32).collect(),
            materials,
            num_curves: curves.count, curve_lut: curves.lut, cie_lut: cie,
            texels, textures, texstack_textures, texstacks,
            imap_row_pdf: row_pdf, imap_row_cdf: row_cdf, imap_marginal_pdf: m_pdf, imap_marginal_cdf: m_cdf,
            cameras: cams, environment: env,
            env_sampling_probability: world.get_env_sampling_probability(), // already 1.0 when there are no lights (world/mod.rs:170-176)
            bounds, unbaked_importance_map: unbaked, env_basis_curves,
        };
        // pointers into the now-pinned vectors
        for (t, tx) in flat.textures.iter_mut().zip(flat.texels.iter()) {
            t.texels = tx.as_ptr();
        }
        if flat.environment.imap_rows > 0 {
            flat.environment.imap_row_pdf = flat.imap_row_pdf.as_ptr();
            flat.environment.imap_row_cdf = flat.imap_row_cdf.as_ptr();
            flat.environment.imap_marginal_pdf = flat.imap_marginal_pdf.as_ptr();
            flat.environment.imap_marginal_cdf = flat.imap_marginal_cdf.as_ptr();
        }
        Ok(flat)
    }

    /// The descriptor; borrows `self` (all pointers point into it).
    pub fn desc(&self) -> RptSceneDesc {
        RptSceneDesc {
            abi_version: RPT_ABI_VERSION,
            num_instances: self.instances.len() as u32, instances: self.instances.as_ptr(),
            num_meshes: self.meshes.len() as u32, meshes: self.meshes.as_ptr(),
            num_lights: self.lights.len() as u32, lights: self.lights.as_ptr(),
            num_materials: self.materials.len() as u32, materials: self.materials.as_ptr(),
            num_curves: self.num_curves as u32, num_lambda: NUM_LAMBDA as u32,
            lut_lambda_lo: self.bounds.lower, lut_lambda_hi: self.bounds.upper,
            curve_lut: self.curve_lut.as_ptr(), cie_lut: self.cie_lut.as_ptr(),
            num_textures: self.textures.len() as u32, textures: self.textures.as_ptr(),
            num_texstack_textures: self.texstack_textures.len() as u32, texstack_textures: self.texstack_textures.as_ptr(),
            num_texstacks: self.texstacks.len() as u32, texstacks: self.texstacks.as_ptr(),
            environment: self.environment,
            env_sampling_probability: self.env_sampling_probability,
            num_cameras: self.cameras.len() as u32, cameras: self.cameras.as_ptr(),
        }
    }

    /// Curve tables of `RptImapBake` for an Unbaked importance map: luminance_curve and every basis curve of the environment's
    /// texture stack at lambda_i = lo + i * (hi - lo) / n (the sample points of Curve::evaluate_integral, importance_map.rs:141-152).
    pub fn imap_bake_tables(&self, luminance: &Curve, n: usize) -> (Vec<f32>, Vec<f32>) {
        let step = (self.bounds.upper - self.bounds.lower) / n as f32;
        let lam = |i: usize| self.bounds.lower + step * i as f32;
        let lum: Vec<f32> = (0..n).map(|i| luminance.evaluate(lam(i))).collect();
        let mut basis = vec![0.0f32; self.env_basis_curves.len() * 4 * n];
        for (k, tex) in self.env_basis_curves.iter().enumerate() {
            for (c, curve) in tex.iter().enumerate() {
                if let Some(curve) = curve {
                    for i in 0..n {
                        basis[(4 * k + c) * n + i] = curve.evaluate(lam(i));
                    }
                }
            }
        }
        (lum, basis)
    }
}

/// The signal of a `Curve::Linear { signal, bounds: (0, 1), mode: Nearest }` (what ImportanceMap::bake_raw builds).
fn linear_signal(c: &Curve) -> Result<&[f32], String> {
    match c {
        Curve::Linear { signal, .. } => Ok(signal.as_slice()),
        _ => Err("importance-map rows are expected to be Curve::Linear (src/world/importance_map.rs:158-176)".into()),
    }
}
