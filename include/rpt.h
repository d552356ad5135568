/*
 * rpt.h — C ABI of the B200 spectral path-tracing backend.
 *
 * This is the drop-in boundary for ONE path of gillett-hernandez/rust-pathtracer:
 * the PT integrator (reference src/integrator/pt.rs) as driven by
 * `render_sampled(integrator, &RenderSettings, &CameraEnum) -> Vec2D<XYZColor>`
 * (reference src/renderer/naive.rs:27-119, src/renderer/tiled.rs:279-542).
 * A host (the reference's Rust `Renderer`, or the Python mirror in
 * rust-pathtracer_b200/) flattens its `World` (reference src/world/mod.rs:18-28)
 * into the plain-pointer structs below, calls rpt_scene_create once, then
 * rpt_render_pt per render setting, and receives the mean CIE XYZ film.
 *
 * Plain C, no torch types, no C++ types. All arrays are host pointers owned by the
 * caller; the library copies what it needs during rpt_scene_create.
 * All functions return 0 on success, non-zero on failure; rpt_last_error() gives
 * the message (thread-local). Nothing panics or throws across this boundary
 * (reference error convention: panics, src/renderer/mod.rs:46-48, pt.rs:577).
 *
 * The same structs are consumed by the CPU oracle (oracle/rpt_oracle.cpp), which is
 * test infrastructure only.
 */
#ifndef RPT_H
#define RPT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPT_ABI_VERSION 8u

/* ---- MaterialId (reference src/materials/mod.rs:22-27) --------------------------
 * Packed as (tag << 16) | table_index; RPT_MAT_NONE = "no override / no id". */
#define RPT_MAT_TAG_MATERIAL 0u
#define RPT_MAT_TAG_LIGHT 1u
#define RPT_MAT_NONE 0xFFFFFFFFu
#define RPT_MAT_PACK(tag, idx) ((((uint32_t)(tag)) << 16) | ((uint32_t)(idx) & 0xFFFFu))
#define RPT_MAT_INDEX(m) ((m) & 0xFFFFu)
#define RPT_MAT_IS_LIGHT(m) ((((m) >> 16) & 0xFFu) == RPT_MAT_TAG_LIGHT)

/* ---- Aggregate kinds (reference src/geometry/mod.rs:17-116) ---------------------- */
enum RptAggregateKind {
  RPT_AGG_RECT = 0,   /* AARect, src/geometry/rect.rs:14-21 */
  RPT_AGG_SPHERE = 1, /* Sphere, src/geometry/sphere.rs:6-10 */
  RPT_AGG_DISK = 2,   /* Disk,   src/geometry/disk.rs:6-11 */
  RPT_AGG_MESH = 3    /* Mesh,   src/geometry/mesh.rs:257-268 */
};

enum RptAxis { RPT_AXIS_X = 0, RPT_AXIS_Y = 1, RPT_AXIS_Z = 2 };

/* Instance (reference src/geometry/instance.rs:9-15). instance_id == array index. */
typedef struct RptInstance {
  uint32_t kind;          /* RptAggregateKind */
  float origin[3];        /* rect / sphere / disk origin (local space) */
  float size[2];          /* rect: size.0,size.1 ; sphere/disk: radius in size[0] */
  uint32_t axis;          /* rect normal axis (RptAxis) */
  uint32_t two_sided;     /* rect / disk */
  int32_t mesh;           /* index into meshes[] for RPT_AGG_MESH, else -1 */
  uint32_t has_transform; /* Option<Transform3> */
  float forward[16];      /* Transform3.forward, row-major 4x4 (local -> world) */
  float reverse[16];      /* Transform3.reverse, row-major 4x4 (world -> local) */
  uint32_t material;      /* Option<MaterialId>: RPT_MAT_PACK(..) or RPT_MAT_NONE */
} RptInstance;

/* Mesh (reference src/geometry/mesh.rs:257-268). Faces are triangles. */
typedef struct RptMesh {
  uint32_t num_vertices;
  uint32_t num_faces;
  const float *vertices;         /* 3*num_vertices */
  const uint32_t *indices;       /* 3*num_faces */
  const float *normals;          /* 3*num_vertices or NULL (no shading normals) */
  const uint32_t *face_material; /* num_faces packed MaterialIds, or NULL => Material(0) */
} RptMesh;

/* ---- Materials (reference src/materials/{lambertian,ggx,diffuse_light,sharp_light}.rs) */
enum RptMaterialType {
  RPT_MATERIAL_LAMBERTIAN = 0,
  RPT_MATERIAL_GGX = 1,
  RPT_MATERIAL_DIFFUSE_LIGHT = 2,
  RPT_MATERIAL_SHARP_LIGHT = 3
};
enum RptSidedness { RPT_SIDED_FORWARD = 0, RPT_SIDED_REVERSE = 1, RPT_SIDED_DUAL = 2 };

typedef struct RptMaterial {
  uint32_t type;      /* RptMaterialType */
  int32_t texstack;   /* Lambertian: index into texstacks[] */
  int32_t curve_a;    /* GGX: eta        | lights: bounce_color */
  int32_t curve_b;    /* GGX: eta_o      | lights: emit_color   */
  int32_t curve_c;    /* GGX: kappa      | unused               */
  float alpha;        /* GGX roughness (ggx.rs:184) */
  float sharpness;    /* SharpLight: already 1+|sharpness| (sharp_light.rs:25) */
  uint32_t sidedness; /* lights: RptSidedness */
  uint32_t metallic;  /* GGX: kappa integral over visible > 0 (ggx.rs:205) */
} RptMaterial;

/* ---- Textures (reference src/texture.rs) ---------------------------------------- */
typedef struct RptTexture {
  uint32_t channels; /* 1 = Texture1, 4 = Texture4 */
  uint32_t width, height;
  const float *texels; /* width*height*channels, row-major y*width+x (vec2d.rs:31-32) */
  int32_t curves[4];   /* curve LUT ids; Texture1 uses curves[0] */
} RptTexture;

typedef struct RptTexStack {
  uint32_t first; /* first index into texstack_textures[] */
  uint32_t count;
} RptTexStack;

/* ---- Environment (reference src/world/environment.rs:7-27) ---------------------- */
enum RptEnvKind { RPT_ENV_CONSTANT = 0, RPT_ENV_SUN = 1, RPT_ENV_HDR = 2 };

typedef struct RptEnvironment {
  uint32_t kind; /* RptEnvKind */
  float strength;
  int32_t curve;           /* Constant / Sun colour LUT id */
  float angular_diameter;  /* Sun */
  float sun_direction[3];  /* Sun, normalised */
  int32_t texstack;        /* HDR texture stack */
  float rot_forward[16];   /* HDR rotation Transform3 */
  float rot_reverse[16];
  /* Baked importance map (reference src/world/importance_map.rs:32-46), or rows==0
   * for Empty/Unbaked (pdf 1/4pi, uniform uv sampling: environment.rs:253,348).
   * Every row is a CurveWithCDF {pdf: Linear Nearest, cdf: Linear Nearest} over [0,1]. */
  uint32_t imap_rows;        /* vertical_resolution  (index <-> u) */
  uint32_t imap_cols;        /* horizontal_resolution (index <-> v) */
  const float *imap_row_pdf; /* rows*cols, normalised per row (importance_map.rs:158-160) */
  const float *imap_row_cdf; /* rows*cols cumulative mass, last == 1 */
  uint32_t imap_marginal_n;  /* length of the marginal tables */
  const float *imap_marginal_pdf; /* marginal pdf signal (rows entries, Nearest) */
  const float *imap_marginal_cdf; /* marginal cdf signal produced by Curve::to_cdf */
  float imap_marginal_integral;   /* CurveWithCDF.pdf_integral of the marginal */
} RptEnvironment;

/* ---- Camera (reference src/camera/projective_camera.rs:8-24, panorama_camera.rs:6-16) -------- */
enum RptCameraKind { RPT_CAMERA_PROJECTIVE = 0, RPT_CAMERA_PANORAMA = 1 };

typedef struct RptCamera {
  float origin[3];
  float u[3], v[3], w[3]; /* projective: w = -direction (projective_camera.rs:44-48); panorama: w = +direction (:31-33) */
  float lower_left[3];    /* projective only */
  float horizontal[3];
  float vertical[3];
  float aperture_diameter; /* projective only; the panorama camera is a pinhole (panorama_camera.rs:69-73) */
  uint32_t kind;           /* RptCameraKind */
  float angle_span[2];     /* panorama: horizontal / vertical field of view in radians (panorama_camera.rs:35-38) */
} RptCamera;

/* ---- Scene: the flattened World (reference src/world/mod.rs:18-28) -------------- */
typedef struct RptSceneDesc {
  uint32_t abi_version; /* RPT_ABI_VERSION */

  uint32_t num_instances;
  const RptInstance *instances;
  uint32_t num_meshes;
  const RptMesh *meshes;
  uint32_t num_lights; /* World.lights: instance ids in push order (world/mod.rs:42-66) */
  const uint32_t *lights;

  uint32_t num_materials; /* index 0 is the mauve error light (parsing/mod.rs:440-444) */
  const RptMaterial *materials;

  /* Curve LUTs: every Curve / CurveWithCDF.pdf the path evaluates, sampled host-side with
   * the real evaluate() on a uniform grid of num_lambda points covering
   * [lut_lambda_lo, lut_lambda_hi] inclusive; the device interpolates linearly. */
  uint32_t num_curves;
  uint32_t num_lambda;
  float lut_lambda_lo, lut_lambda_hi;
  const float *curve_lut; /* num_curves * num_lambda */
  const float *cie_lut;   /* 3 * num_lambda: x_bar, y_bar, z_bar on the same grid */

  uint32_t num_textures;
  const RptTexture *textures;
  uint32_t num_texstack_textures;
  const uint32_t *texstack_textures; /* texture ids, concatenated per stack */
  uint32_t num_texstacks;
  const RptTexStack *texstacks;

  RptEnvironment environment;
  float env_sampling_probability; /* scene value (parsing/mod.rs:559); forced to 1 by the
                                     library when num_lights == 0 (world/mod.rs:77-80) */
  uint32_t num_cameras;
  const RptCamera *cameras; /* already aspect-corrected (parsing/cameras.rs:191-201) */
} RptSceneDesc;

/* ---- Render parameters: RenderSettings + IntegratorKind::PT ----------------------
 * reference src/parsing/config.rs:45-62,21-24; src/integrator/mod.rs:59-105 */
typedef struct RptRenderParams {
  uint32_t width, height;
  uint32_t spp;        /* samples rendered by THIS call (this rank's share) */
  uint32_t spp_offset; /* global index of this call's first sample (multi-GPU split) */
  uint32_t spp_total;  /* divisor for the mean (min_samples); 0 => leave the SUM in film */
  uint32_t min_bounces; /* russian-roulette start index (pt.rs:475) */
  uint32_t max_bounces;
  uint32_t light_samples;
  uint32_t only_direct;
  float lambda_lo, lambda_hi; /* wavelength_bounds */
  uint32_t camera;            /* index into cameras[] */
  uint64_t seed;              /* Philox key */
  uint32_t flags;             /* RPT_FLAG_*: run-time instrumentation, off by default */
  uint32_t reserved;          /* 0 */
} RptRenderParams;

/* RptRenderParams.flags. Instrumentation is opt-in: a plain render records two CUDA events and carries no counters. */
#define RPT_FLAG_KERNEL_TIMES 1u /* CUDA events around every kernel launch -> rpt_last_kernel_times() */
#define RPT_FLAG_BVH_STATS 2u    /* nodes / triangles / instances visited -> RptCounters.walk_* / shadow_* (else 0) */

/* Profile counters (reference src/profile.rs:1-8) + true BVH-query counts. */
typedef struct RptCounters {
  uint64_t camera_rays;
  uint64_t bounce_rays; /* path vertices incl. the camera vertex (integrator/utils.rs:375) */
  uint64_t shadow_rays;
  uint64_t light_rays;
  uint64_t env_hits;
  uint64_t segments;     /* walk rays traced (one iteration of integrator/utils.rs:170) */
  uint64_t true_rays;    /* closest-hit queries issued: walk + NEE */
  uint64_t kernel_launches; /* CUDA kernels launched by this call */
  /* BVH work of the two traversal kernels, for the roofline accounting:
   * bytes fetched = nodes * 64 + triangles * 48 + instances * 144 */
  uint64_t shadow_rays_traced; /* NEE rays actually traced (zero-weight ones are skipped) */
  uint64_t walk_nodes, walk_tris, walk_insts;
  uint64_t shadow_nodes, shadow_tris, shadow_insts;
  uint64_t nee_vertices; /* path vertices that ran NEE (non-light surface vertices when light_samples > 0) */
  double device_ms; /* device time of the call: first to last CUDA event on the library's stream */
} RptCounters;

typedef struct RptScene RptScene; /* opaque */

/* Per-kernel device time of the most recent rpt_render_* call (CUDA events on the
 * library's stream). Kernel names are stable identifiers used by bench.py / profiles. */
typedef struct RptKernelTime {
  const char *name;
  uint32_t launches;
  float ms;
} RptKernelTime;

const char *rpt_last_error(void);
uint32_t rpt_abi_version(void);
int rpt_device_count(int *count);

/* Flatten + upload. `device` is the CUDA ordinal this scene lives on. Replaces:
 * Integrator::from_settings_and_world + Arc<World> (tiled.rs:553-651). */
int rpt_scene_create(const RptSceneDesc *desc, int device, RptScene **out);
int rpt_scene_destroy(RptScene *scene);

/* Replaces render_sampled() (tiled.rs:279-542 / naive.rs:27-119) for PathTracingIntegrator.
 * film_xyzw: HOST buffer, width*height*4 floats, row-major y*width+x, w lane = 0.
 * The call copies the film device -> host. Mean XYZ when spp_total > 0 (tiled.rs:396-398). */
int rpt_render_pt(RptScene *scene, const RptRenderParams *params, float *film_xyzw,
                  RptCounters *counters);

/* Same render, film left on the device. *film_dev receives a device pointer owned by the
 * scene (valid until the next render call or destroy): width*height float4, the
 * UN-NORMALISED sum when spp_total == 0. Used for the multi-GPU NCCL reduce and for
 * device-resident timing. */
int rpt_render_pt_device(RptScene *scene, const RptRenderParams *params, void **film_dev,
                         RptCounters *counters);

/* Parity hook (a): closest hit of the primary ray through every pixel centre-jittered by
 * sample 0 (same Philox draws as the render). Host outputs, width*height each.
 * instance id / primitive id (triangle index, 0 for analytic shapes) / t; miss = 0xFFFFFFFF. */
int rpt_trace_primary(RptScene *scene, const RptRenderParams *params, uint32_t *instance_id,
                      uint32_t *primitive_id, float *t);

/* Generic closest-hit batch query: n rays (origin xyz, dir xyz, tmax) from host arrays. */
int rpt_trace_rays(RptScene *scene, uint32_t n, const float *origins, const float *dirs,
                   const float *tmax, uint32_t *instance_id, uint32_t *primitive_id, float *t);

/* Device-side film normalisation for callers that reduced sums themselves (NCCL):
 * film[i] *= scale. film_dev is any device pointer of n_float4 float4s on the scene's device. */
int rpt_film_scale(RptScene *scene, void *film_dev, uint64_t n_float4, float scale);

/* Timing of the last render: fills up to cap entries, returns count in *n. */
int rpt_last_kernel_times(RptScene *scene, RptKernelTime *out, uint32_t cap, uint32_t *n);

/* ---- output_film: the step right after the hot path -------------------------------------
 * reference src/renderer/mod.rs:24-80 (output_film) -> src/tonemap/{clamp,reinhard0,reinhard1}.rs (initialize + map)
 * -> src/tonemap/mod.rs:207-338 (write_to_files): linear RGB in the chosen primaries for the EXR, tonemapped +
 * OETF-encoded 8-bit RGBA for the PNG. File encoding itself stays on the host. */
enum RptTonemapper { RPT_TONEMAP_CLAMP = 0, RPT_TONEMAP_REINHARD0 = 1, RPT_TONEMAP_REINHARD1 = 2 };
enum RptColorSpace { RPT_COLORSPACE_SRGB = 0, RPT_COLORSPACE_REC709 = 1, RPT_COLORSPACE_REC2020 = 2 };
typedef struct RptOutputSettings {
  uint32_t tonemapper;     /* RptTonemapper (parsing/tonemap.rs:9-31) */
  uint32_t luminance_only; /* false selects the per-channel x3 variants of Reinhard0/1 */
  float exposure;          /* Clamp: 2^exposure multiplier (clamp.rs:83) */
  float key_value;         /* Reinhard0/1 */
  float white_point;       /* Reinhard1 */
  uint32_t colorspace;     /* RptColorSpace (parsing/config.rs:33-43) */
  float factor;            /* output_film's factor x premultiply (renderer/mod.rs:25) */
} RptOutputSettings;

/* Tonemaps the film. film_xyzw: host film (w*h*4 floats) or NULL to use the device-resident film of the scene's
 * last render (no re-upload). rgb_linear: optional host out, w*h*3 floats = M(factor * XYZ) (the EXR payload).
 * rgba8: host out, w*h*4 bytes = ceil(255 * OETF(M(map(XYZ)))) clamped, alpha 255 (the PNG payload).
 * l_w: optional out, the 4 lanes of the tonemapper's log-average (lane 0..2 = X,Y,Z for x3 variants; Y only else). */
int rpt_output_film(RptScene *scene, const float *film_xyzw, uint32_t width, uint32_t height, const RptOutputSettings *settings,
                    float *rgb_linear, uint8_t *rgba8, float *l_w);

/* ---- N3: ImportanceMap::bake_raw on the device (reference src/world/importance_map.rs:78-253) --------------------
 * The step immediately before the hot path: for every (row, column) of the map, the luminance of the environment texel
 * at uv = (row / rows, column / cols) (:137-140) is the num_samples-point left Riemann sum over the wavelength bounds of
 *   Machine{1, [Mul luminance_curve, Mul texture_stack.curve_at(uv)]}          (:141-152; texture.rs:41-77,126-131,221-228)
 * every row is normalised into a pdf and a cumulative mass function (:153-176), the row sums into the marginal (:214),
 * and the marginal goes through Curve::to_cdf((0,1), 100) (:239-244).
 * The curves are evaluated by the CALLER (the real math::Curve on the host side of the shim) at the sample wavelengths
 * lambda_i = lo + i * (hi - lo) / num_samples, i < num_samples, so no curve interpolation happens on the device. */
typedef struct RptImapBake {
  uint32_t rows, cols;        /* vertical_resolution (index <-> u), horizontal_resolution (index <-> v) */
  uint32_t num_samples;       /* num_samples_for_texel_spectra, 100 in the reference (:85) */
  float lambda_lo, lambda_hi; /* wavelength_bounds */
  const float *luminance;     /* num_samples values of luminance_curve */
  const float *basis;         /* environment texture stack, texture k, channel c: basis[(4 * k + c) * num_samples + i]
                                 = curves[c].pdf evaluated at lambda_i (Texture1 uses c = 0 only) */
} RptImapBake;

/* Bakes from the scene's device-resident environment texels and INSTALLS the tables in the scene (subsequent renders
 * importance-sample the environment through them; nothing is re-uploaded). Every out pointer is optional (NULL = keep on
 * the device only): row_pdf / row_cdf rows*cols floats, marginal_pdf / marginal_cdf rows floats, marginal_integral 1 float
 * — exactly the payload of the reference's on-disk cache (:254-324), whose bincode encoding stays on the host. */
int rpt_scene_bake_importance_map(RptScene *scene, const RptImapBake *bake, float *row_pdf, float *row_cdf, float *marginal_pdf,
                                  float *marginal_cdf, float *marginal_integral);

/* ---- Multi-GPU: the spp split + film exchange behind the C ABI -----------------------------------------
 * The reference is one process with one `Renderer::render` call (src/bin/main.rs:59-68,170; src/renderer/mod.rs:107-112),
 * so a `CudaRenderer { devices: Vec<u32> }` needs the whole multi-GPU job behind ONE call from ONE host thread.
 * rpt_multi_create uploads a replica of the scene to every listed device (devices[0] is the root) and prepares the film
 * exchange; rpt_multi_render_pt splits params->spp over the devices (remainder to the low ranks; device i continues the
 * Philox sample index where device i-1 stopped, so the samples are exactly those of a single-device render of params->spp),
 * renders with one host worker thread per device, sums the un-normalised XYZ films onto the root, normalises by
 * params->spp_total and downloads the film (film_xyzw may be NULL: the result stays on the root device,
 * rpt_multi_scene(multi, 0) + rpt_output_film can tonemap it there).
 * Exchange: RPT_MULTI_PEER = one fused reduce-scatter + normalise + gather kernel per device over NVLink peer mappings
 * (default when the devices can map each other); RPT_MULTI_NCCL = ncclReduce(sum) to the root + normalisation kernel
 * (NCCL is dlopen'ed on first use: libnccl.so.2, or $RPT_NCCL_LIB; forced with RPT_MULTI_REDUCE=nccl). */
enum RptMultiMethod { RPT_MULTI_PEER = 0, RPT_MULTI_NCCL = 1 };
typedef struct RptMulti RptMulti; /* opaque */
typedef struct RptMultiTimes {
  uint32_t method;               /* RptMultiMethod actually used */
  uint32_t devices;
  double render_device_ms_max;   /* slowest device's render, CUDA events on its stream */
  double exchange_device_ms;     /* film exchange + normalisation, CUDA events, max over devices */
  double render_wall_ms;         /* host wall clock of the three phases of the call */
  double exchange_wall_ms;
  double download_wall_ms;
} RptMultiTimes;
int rpt_multi_create(const RptSceneDesc *desc, const int *devices, int n, RptMulti **out);
int rpt_multi_destroy(RptMulti *multi);
/* The replica on devices[index] (for rpt_output_film, rpt_scene_stats, rpt_last_kernel_times ...). Owned by `multi`. */
int rpt_multi_scene(RptMulti *multi, int index, RptScene **scene);
/* rpt_scene_bake_importance_map on every replica (tables stay on the devices). */
int rpt_multi_bake_importance_map(RptMulti *multi, const RptImapBake *bake);
/* counters: sums over the devices; counters->device_ms = slowest device's render + exchange. times may be NULL. */
int rpt_multi_render_pt(RptMulti *multi, const RptRenderParams *params, float *film_xyzw, RptCounters *counters, RptMultiTimes *times);
/* One-shot form: create, render, destroy. */
int rpt_render_pt_multi(const RptSceneDesc *desc, const int *devices, int n, const RptRenderParams *params, float *film_xyzw,
                        RptCounters *counters);

/* Bandwidth probe: the measured denominators of the roofline fractions (north_star: "achieved HBM/L2 GB/s ... against B200
 * peak"; SURVEY §8d asks for an L2-resident streaming-kernel peak measured on the box). Allocates `bytes`, runs the pattern
 * `reps` times per launch, returns the best of 4 timed launches in GB/s. mode 0: streaming 128-bit reads (L2 bandwidth when
 * bytes fits L2, HBM read bandwidth when far larger); mode 1: copy (read + write); mode 2: random 64-byte gathers (the BVH
 * node fetch pattern). Measurement infrastructure; not used by any render call. */
int rpt_probe_bandwidth(int device, uint64_t bytes, uint32_t reps, int mode, double *gbps);

/* BVH statistics for roofline accounting: bytes of nodes / primitives. */
typedef struct RptSceneStats {
  uint64_t tlas_nodes, blas_nodes, triangles, instances;
  uint64_t node_bytes, triangle_bytes, scene_bytes_total;
} RptSceneStats;
int rpt_scene_stats(RptScene *scene, RptSceneStats *out);

#ifdef __cplusplus
}
#endif
#endif /* RPT_H */
