// rpt_device.cuh — device-side scene layout and per-ray functions of the PT path (sm_100a).
//
// Every function cites the reference file:line (gillett-hernandez/rust-pathtracer) whose
// behaviour it reproduces. This is the product path: it never touches oracle/.
//
// Numerics: fp32 throughout like the reference; IEEE div/sqrt (no --use_fast_math); the library is
// compiled with -fmad=false because Rust never contracts a*b+c into an FMA: a fused cross product turns
// the exact 0 of an axis-aligned normal into 1e-9, which flips the tangent frame's copysign branch and
// sends the path elsewhere. The hit/miss DECISION arithmetic of the ray/primitive tests additionally
// spells its roundings out with __f*_rn intrinsics (parity check (a) of BASELINE.json). Explicit fmaf()
// is used only in the BVH slab test, which prunes but never decides.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rpt.h"
#include "../../include/rpt_rng.h"

#define RPT_PI 3.14159265358979323846f
#define RPT_TAU 6.28318530717958647692f
#define RPT_EPS 1.1920929e-7f
#define RPT_NORMAL_OFFSET 0.001f /* reference src/lib.rs:48 */
#define RPT_NONE 0xFFFFFFFFu
#define RPT_INF __int_as_float(0x7f800000)
// BVH work counters (nodes / triangles / instances visited, reported through RptCounters) are a RUN-TIME opt-in
// (RptRenderParams.flags & RPT_FLAG_BVH_STATS): the traversal code is instantiated with and without them (template
// parameter STATS), so a plain render carries no instrumentation.
#define RPT_STAT(x)  \
  do {               \
    if (STATS) { x; } \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Device scene (SoA buffers in HBM; pointers passed by value inside DevScene as a kernel param)
// ---------------------------------------------------------------------------------------------

// BVH2 node, 64 B = 4 x float4 (one 128-bit load each). Both children's boxes live in the parent,
// so one node fetch decides both descents. child >= 0: inner node index; child < 0: leaf, ~child =
// shape index (instance id in the TLAS, mesh-local triangle id in a BLAS).
struct __align__(16) DevNode {
  float4 lmin_lmaxx;  // l.min.xyz, l.max.x
  float4 lmaxyz_rminxy;  // l.max.y, l.max.z, r.min.x, r.min.y
  float4 rminz_rmax;  // r.min.z, r.max.xyz
  int4 children;      // left, right, unused, unused
};

enum : uint32_t {
  DI_KIND_MASK = 0x3u,
  DI_AXIS_SHIFT = 2,        // 2 bits
  DI_TWO_SIDED = 1u << 4,
  DI_HAS_TRANSFORM = 1u << 5,
};

// Instance record, 144 B. rev/fwd are the upper 3x4 of Transform3.reverse / .forward.
struct __align__(16) DevInstance {
  float4 rev[3];
  float4 fwd[3];
  float4 origin_size0;  // origin.xyz, size[0] (or radius)
  float size1;
  uint32_t flags;     // kind | axis << 2 | two_sided | has_transform
  uint32_t material;  // packed MaterialId override or RPT_NONE
  int32_t blas_root;  // child-ref of the mesh's BLAS root (absolute node index, or ~tri for 1-triangle meshes)
  uint32_t tri_base;  // first triangle of the mesh in the global triangle arrays
  uint32_t order;     // position in the reference's flat-BVH candidate order (tie-break only)
  uint32_t has_normals;
  int32_t blas_root4;  // the same root in the four-wide tree (DevScene::nodes4)
};

struct DevTexture {
  const float *texels;
  uint32_t channels, width, height;
  int32_t curves[4];
};

struct DevScene {
  // acceleration structure
  const DevNode *nodes;  // TLAS nodes first, then every BLAS
  // Four-wide form of the same trees (RPT_BVH4=1; nullptr otherwise): 8 float4 = 128 B per node, see TravT<true>.
  const float4 *nodes4;
  int32_t tlas_root4;  // child-ref into nodes4 (leaf refs are the same as in the two-wide tree)
  // TLAS leaf table: .x = instance id, .y = mesh-local triangle id or RPT_NONE (whole instance),
  // .z = global triangle index, .w = the instance's candidate order. Untransformed mesh instances are
  // flattened into the TLAS triangle by triangle (exact: their local space IS world space), so the common
  // "one big untransformed mesh" scene traverses a single-level tree.
  const uint4 *tlas_leaves;
  int32_t tlas_root;     // child-ref (leaf refs index tlas_leaves)
  uint32_t num_instances;
  const DevInstance *instances;
  const float4 *tri_verts;    // 3 float4 per triangle: p0|material, p1|order, p2|unused  (w lanes as bits)
  const float4 *tri_normals;  // 3 float4 per triangle (only for meshes with shading normals)
  // lights / materials / spectra
  uint32_t num_lights;
  const uint32_t *lights;
  // Every instance whose hits carry a Light material, when all of them are analytic shapes (and there are few):
  // lets the NEE visibility query run as "closest light, then any occluder in front of it" (k_shadow).
  uint32_t min_grab;        // smallest number of queue tiles a warp claims at once (TileStream)
  uint32_t num_light_geom;  // 0 = two-phase NEE visibility disabled
  const uint32_t *light_geom;
  const float *light_geom_box;  // 6 floats (world min, max) per entry: the box the reference's BVH gates the shape with
  const RptMaterial *materials;
  // Per material: (curve LUT id as bits, texel) when its texture stack is exactly one Texture1 of one texel (every
  // `lambertian_*` of data/lib_materials.toml is: single_pixel.png x a reflectance curve), else (-1, 0). Lets the shade
  // kernel skip four dependent loads (stack -> texture ids -> texture record -> texel) on its critical path; the value is
  // the one texstack_eval returns (0 + curve * texel).
  const float2 *mat_fast;
  const float *curve_lut;
  const float *cie_lut;
  uint32_t num_lambda;
  float lut_lo, lut_hi;
  const DevTexture *textures;
  const uint32_t *stack_tex;
  const RptTexStack *stacks;
  // environment
  uint32_t env_kind;
  float env_strength;
  int32_t env_curve;
  float env_angular_diameter;
  float3 env_sun_dir;
  int32_t env_texstack;
  float4 env_rot_fwd[3], env_rot_rev[3];
  uint32_t env_unrotated;  // both are the identity: the uv -> direction -> uv round trips take uv_roundtrip_unrotated_cr
  uint32_t imap_rows, imap_cols, imap_marginal_n;
  const float *imap_row_pdf, *imap_row_cdf, *imap_m_pdf, *imap_m_cdf;
  float imap_marginal_integral;
  // Guide tables of the CDF inversions (nullptr = plain binary search): per CDF, RPT_IMAP_GUIDE entries; entry k is the first
  // index whose CDF value reaches k / RPT_IMAP_GUIDE of the CDF's top, so an inversion searches a few entries instead of the row.
  const uint32_t *imap_row_guide, *imap_m_guide;
  // Small-scene mode (<= RPT_SMALL_MAX leaves, no BLAS): the traversal kernels skip the BVH and test every leaf for every
  // ray in a warp-uniform loop out of shared memory (32 of 32 lanes busy, no stack); see SmallTrav below.
  uint32_t small_n;          // 0 = mode off; else the number of leaves
  uint32_t small_ntri;       // leaves [0, small_ntri) are triangles, [small_ntri, small_n) whole analytic instances
  const float4 *small_tris;  // per triangle leaf 9 float4: the three vertices in each of the 3 watertight axis permutations
  const uint4 *small_leaves; // per leaf: instance id, mesh-local triangle id, triangle candidate order, instance candidate order
  const float *small_boxes;  // per instance leaf (index k - small_ntri) 6 floats: the box the reference's BVH gates the shape with
  float3 world_center; // centre of the scene bounds (ray binning only)
  float3 world_min, world_inv_extent;  // scene bounds for the NEE origin-cell grid (ray binning only)
  float p_env;         // effective env sampling probability (1 when there are no lights)
  float world_radius;  // World.radius (world/mod.rs:69-72); informational
};

// ---------------------------------------------------------------------------------------------
// float3 helpers (math::Vec3 / Point3)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// dot product with the reference's rounding: three products, two adds, nothing fused
__device__ __forceinline__ float dot_rn(float3 a, float3 b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ float norm_squared(float3 a) { return dot(a, a); }
__device__ __forceinline__ float3 normalized(float3 a) { return a / sqrtf(dot(a, a)); }  // Vec3::normalized = v / norm
__device__ __forceinline__ float signumf(float x) { return (x != x) ? x : copysignf(1.0f, x); }  // f32::signum
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

__device__ __forceinline__ float3 xform_point(const float4 *m, float3 p) {
  return f3(m[0].x * p.x + m[0].y * p.y + m[0].z * p.z + m[0].w, m[1].x * p.x + m[1].y * p.y + m[1].z * p.z + m[1].w,
            m[2].x * p.x + m[2].y * p.y + m[2].z * p.z + m[2].w);
}
__device__ __forceinline__ float3 xform_vec(const float4 *m, float3 v) {
  return f3(m[0].x * v.x + m[0].y * v.y + m[0].z * v.z, m[1].x * v.x + m[1].y * v.y + m[1].z * v.z,
            m[2].x * v.x + m[2].y * v.y + m[2].z * v.z);
}
__device__ __forceinline__ float3 xform_vec_transposed(const float4 *m, float3 v) {  // (M^T) v, 3x3 part
  return f3(m[0].x * v.x + m[1].x * v.y + m[2].x * v.z, m[0].y * v.x + m[1].y * v.y + m[2].y * v.z,
            m[0].z * v.x + m[1].z * v.y + m[2].z * v.z);
}

// math::TangentFrame::from_normal (Duff et al. 2017)
struct Frame {
  float3 t, b, n;
};
__device__ __forceinline__ Frame frame_from_normal(float3 n) {
  float sign = copysignf(1.0f, n.z);
  float a = -1.0f / (sign + n.z);
  float b = n.x * n.y * a;
  Frame f;
  f.t = f3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
  f.b = f3(b, sign + n.y * n.y * a, -n.y);
  f.n = n;
  return f;
}
__device__ __forceinline__ float3 to_local(const Frame &f, float3 v) { return f3(dot(f.t, v), dot(f.b, v), dot(f.n, v)); }
__device__ __forceinline__ float3 to_world(const Frame &f, float3 v) { return f.t * v.x + f.b * v.y + f.n * v.z; }

// math sampling helpers
__device__ __forceinline__ float3 random_cosine_direction(float sx, float sy) {
  float s, c;
  sincosf(RPT_TAU * sx, &s, &c);
  float r = sqrtf(sy);
  return f3(c * r, s * r, sqrtf(1.0f - sy));
}
__device__ __forceinline__ float3 random_on_unit_sphere(float sx, float sy) {
  float s, c;
  sincosf(sx * RPT_TAU, &s, &c);
  float z = sy * 2.0f - 1.0f;
  float r = sqrtf(1.0f - z * z);
  return f3(r * c, r * s, z);
}
__device__ __forceinline__ float3 random_in_unit_disk(float sx, float sy) {
  float s, c;
  sincosf(sx * RPT_TAU, &s, &c);
  float v = sqrtf(sy);
  return f3(c * v, s * v, 0.0f);
}
__device__ __forceinline__ float3 uv_to_direction(float u, float v) {
  float st, ct, sp, cp;
  sincosf((u - 0.5f) * RPT_TAU, &st, &ct);
  sincosf(v * RPT_PI, &sp, &cp);
  return f3(sp * ct, sp * st, cp);
}
__device__ __forceinline__ void direction_to_uv(float3 d, float &u, float &v) {
  float theta = atan2f(d.y, d.x);
  float phi = acosf(d.z);
  u = theta / 2.0f / RPT_PI + 0.5f;
  v = phi / RPT_PI;
}
// The same maps with their trigonometry evaluated as CORRECTLY ROUNDED f32 (through f64), used by the HDR environment only:
// importance-map samples sit exactly on texel boundaries (nearest-mode CDF inversion returns grid points) and then go
// through uv -> direction -> rotate -> uv -> texel, so a 1-ulp difference between two libm implementations flips the texel
// that is read. The reference inherits whatever the platform libm does there; both this file and the oracle pin it to the
// correctly rounded value. Out of line: the f64 code must not cost the other paths registers.
__device__ __noinline__ float3 uv_to_direction_cr(float u, float v) {
  double st, ct, sp, cp;
  sincos((double)((u - 0.5f) * RPT_TAU), &st, &ct);
  sincos((double)(v * RPT_PI), &sp, &cp);
  float fst = (float)st, fct = (float)ct, fsp = (float)sp, fcp = (float)cp;
  return f3(fsp * fct, fsp * fst, fcp);
}
__device__ __noinline__ float2 direction_to_uv_cr(float3 d) {
  float theta = (float)atan2((double)d.y, (double)d.x);
  float phi = (float)acos((double)d.z);
  return make_float2(theta / 2.0f / RPT_PI + 0.5f, phi / RPT_PI);
}
// direction_to_uv_cr(uv_to_direction_cr(u, v)) — the round trip every HDR-environment lookup makes (environment.rs:56-98,
// 198-258, 303-353) — for an UNROTATED environment, without the f64 atan2 / acos (27 % of the environment NEE kernel's
// instructions on the GGX + HDR scene). The direction is the f32-rounded image of two angles whose f64 sines and cosines are
// already at hand, so each angle of the way back is the angle of the way out plus a tiny rotation that needs no inverse
// trigonometry:
//   atan2(dy, dx) = a + atan2(dy cos a - dx sin a, dx cos a + dy sin a)         (|.| ~ 1e-7: atan t = t - t^3 / 3)
//   acos(dz)      = b + atan2(w cos b - dz sin b, dz cos b + w sin b), w = sqrt((1 - dz)(1 + dz))
// evaluated in f64 (absolute error ~3e-16, that of the libm calls it replaces), then rounded to f32 once like they are.
// Lanes where the rotation is not tiny (poles: dz = +-1 or v within ~1e-4 of them; a degenerate azimuth) take the libm path.
// tests/test_host_logic.py checks the identity against libm on 1.4e8 inputs (0 differing f32 results);
// tests/test_gpu_parity.py::test_env_roundtrip_fast_path compares the two device paths through rpt_debug_env_roundtrip.
__device__ __noinline__ float2 uv_roundtrip_unrotated_cr(float u, float v) {
  const float a = (u - 0.5f) * RPT_TAU, b = v * RPT_PI;  // uv_to_direction_cr's own f32 angles
  // (the polar angle first, then the azimuth: at most one pair of f64 sine / cosine is live at a time)
  float phi, fsp, fcp;
  {
    double sp, cp;
    sincos((double)b, &sp, &cp);
    fsp = (float)sp;
    fcp = (float)cp;
    const double z = (double)fcp;  // direction.z
    const double w = sqrt((1.0 - z) * (1.0 + z));
    const double c2 = fma(z, cp, w * sp), s2 = fma(w, cp, -(z * sp));
    const double t = s2 * (2.0 - c2);  // 1 / c2 = 2 - c2 + O((1 - c2)^2), c2 = cos(rotation) = 1 - O(1e-7)
    if (fabs(z) != 1.0 && fabs(t) < 1e-3)
      phi = (float)((double)b + t * (1.0 - t * t * (1.0 / 3.0)));
    else
      phi = (float)acos(z);
  }
  float theta;
  {
    double st, ct;
    sincos((double)a, &st, &ct);
    const float dx = fsp * (float)ct, dy = fsp * (float)st;  // direction.x, direction.y
    const double x = (double)dx, y = (double)dy;
    const double c = fma(x, ct, y * st), s = fma(y, ct, -(x * st));
    const double t = s / (c > 1e-30 ? c : 1.0);
    if (c > 1e-30 && fabs(t) < 1e-3) {
      const double kPi = 3.14159265358979323846;
      double th = (double)a + t * (1.0 - t * t * (1.0 / 3.0));
      if (th > kPi) th -= 2.0 * kPi;
      else if (th < -kPi) th += 2.0 * kPi;
      theta = (float)th;
    } else {
      // (the poles.) The general path multiplies by the identity matrix in f32, which changes nothing but the sign of a zero
      // component (-0 + 0 = +0) - and that sign is what atan2 decides by here.
      const float ndx = 1.0f * dx + 0.0f * dy + 0.0f * fcp, ndy = 0.0f * dx + 1.0f * dy + 0.0f * fcp;  // xform_vec(identity, d)
      theta = (float)atan2((double)ndy, (double)ndx);
    }
  }
  return make_float2(theta / 2.0f / RPT_PI + 0.5f, phi / RPT_PI);
}
__device__ __forceinline__ float power_heuristic(float a, float b) { return (a * a) / (a * a + b * b); }
__device__ __forceinline__ float power_heuristic_generic(float a, float b) { return a / (a + b); }  // src/lib.rs:114-119
__device__ __forceinline__ bool choose(float x, float split, float &rescaled) {  // Sample1D::choose
  if (x < split) {
    rescaled = clampf(x / split, 0.0f, 1.0f - RPT_EPS);
    return true;
  }
  rescaled = clampf((x - split) / (1.0f - split), 0.0f, 1.0f - RPT_EPS);
  return false;
}

// ---------------------------------------------------------------------------------------------
// spectral LUTs (include/rpt.h: uniform grid, linear interpolation)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lut_eval(const float *__restrict__ lut, uint32_t n, float lo, float hi, float lambda) {
  float x = (lambda - lo) / (hi - lo) * (float)(n - 1);
  x = clampf(x, 0.0f, (float)(n - 1));
  uint32_t i = (uint32_t)x;
  if (i > n - 2) i = n - 2;
  float t = x - (float)i;
  float a = __ldg(lut + i), b = __ldg(lut + i + 1);
  return __fadd_rn(a, __fmul_rn(t, __fsub_rn(b, a)));
}
__device__ __forceinline__ float curve_eval(const DevScene &S, int32_t c, float lambda) {
  return lut_eval(S.curve_lut + (size_t)c * S.num_lambda, S.num_lambda, S.lut_lo, S.lut_hi, lambda);
}
// Texture{1,4}::eval_at + TexStack::eval_at (texture.rs:101-116,134-142,258-266; vec2d.rs:34-42)
__device__ __forceinline__ float texstack_eval(const DevScene &S, int32_t stack, float lambda, float u, float v) {
  float energy = 0.0f;
  RptTexStack st = S.stacks[stack];
  for (uint32_t k = 0; k < st.count; ++k) {
    const DevTexture &T = S.textures[S.stack_tex[st.first + k]];
    float uu = clampf(u, 0.0f, 1.0f - RPT_EPS), vv = clampf(v, 0.0f, 1.0f - RPT_EPS);
    size_t x = (size_t)(uu * (float)T.width), y = (size_t)(vv * (float)T.height);
    const float *tx = T.texels + (y * T.width + x) * T.channels;
    if (T.channels == 1) {
      energy += curve_eval(S, T.curves[0], lambda) * __ldg(tx);
    } else {
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < 4; ++c) s += curve_eval(S, T.curves[c], lambda) * __ldg(tx + c);
      energy += s;
    }
  }
  return energy;
}

// ---------------------------------------------------------------------------------------------
// ray / primitive tests
// ---------------------------------------------------------------------------------------------
struct LocalHit {  // reference HitRecord fields an aggregate produces (hittable.rs:7-16)
  float t;
  float3 p, n;
  float u, v;
  uint32_t material;
};

__device__ __forceinline__ float3 shuffle_axis(float3 v, uint32_t axis) {  // rect.rs:6-12
  return axis == RPT_AXIS_X ? f3(v.z, v.y, v.x) : (axis == RPT_AXIS_Y ? f3(v.x, v.z, v.y) : v);
}
__device__ __forceinline__ float3 axis_vec(uint32_t axis) {
  return axis == RPT_AXIS_X ? f3(1, 0, 0) : (axis == RPT_AXIS_Y ? f3(0, 1, 0) : f3(0, 0, 1));
}

// AARect::hit (rect.rs:69-112). Returns t only (decision part); rect_finish fills the record.
__device__ __forceinline__ bool rect_test(const DevInstance &I, float3 o, float3 d, float t0, float t1, float tmax, float &t_out) {
  uint32_t axis = (I.flags >> DI_AXIS_SHIFT) & 3u;
  float3 org = f3(I.origin_size0);
  float3 tmp_o = shuffle_axis(o - org, axis);
  float3 tmp_d = shuffle_axis(d, axis);
  if (tmp_d.z == 0.0f) return false;
  float t = (-tmp_o.z) / tmp_d.z;
  if (t <= t0 || t > t1 || t >= tmax) return false;
  float xh = __fadd_rn(tmp_o.x, __fmul_rn(t, tmp_d.x)), yh = __fadd_rn(tmp_o.y, __fmul_rn(t, tmp_d.y));
  float hx = I.origin_size0.w / 2.0f, hy = I.size1 / 2.0f;
  if (xh < -hx || xh > hx || yh < -hy || yh > hy) return false;
  t_out = t;
  return true;
}
__device__ __forceinline__ void rect_finish(const DevInstance &I, float3 o, float3 d, float t, LocalHit &h) {
  uint32_t axis = (I.flags >> DI_AXIS_SHIFT) & 3u;
  float3 org = f3(I.origin_size0);
  float3 tmp_o = shuffle_axis(o - org, axis);
  float3 tmp_d = shuffle_axis(d, axis);
  float xh = __fadd_rn(tmp_o.x, __fmul_rn(t, tmp_d.x)), yh = __fadd_rn(tmp_o.y, __fmul_rn(t, tmp_d.y));
  float hx = I.origin_size0.w / 2.0f, hy = I.size1 / 2.0f;
  float3 n = axis_vec(axis);
  if ((I.flags & DI_TWO_SIDED) && dot(d, n) > 0.0f) n = -n;
  h.t = t;
  h.p = o + d * t;
  h.u = (xh + hx) / I.origin_size0.w;
  h.v = (yh + hy) / I.size1;
  h.n = n;  // unit axis; HitRecord::new's normalisation is the identity here
  h.material = RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
}
// Sphere::hit (sphere.rs:34-87)
__device__ __forceinline__ bool sphere_test(const DevInstance &I, float3 o, float3 d, float t0, float t1, float tmax, float &t_out) {
  float radius = I.origin_size0.w;
  float3 oc = o - f3(I.origin_size0);
  // b*b - a*c cancels catastrophically near the silhouette: keep every rounding of the reference
  float a = dot_rn(d, d), b = dot_rn(oc, d), c = __fsub_rn(dot_rn(oc, oc), __fmul_rn(radius, radius));
  float disc = __fsub_rn(__fmul_rn(b, b), __fmul_rn(a, c));
  if (!(disc > 0.0f)) return false;
  float ds = sqrtf(disc);
  float time = (-b - ds) / a;
  if (time < t1 && time > t0 && time < tmax) {
    t_out = time;
    return true;
  }
  time = (-b + ds) / a;
  if (time < t1 && time > t0 && time < tmax) {
    t_out = time;
    return true;
  }
  return false;
}
__device__ __forceinline__ void sphere_finish(const DevInstance &I, float3 o, float3 d, float t, LocalHit &h) {
  float3 p = o + d * t;
  h.t = t;
  h.p = p;
  h.u = h.v = 0.0f;
  h.n = normalized((p - f3(I.origin_size0)) / I.origin_size0.w);
  h.material = RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
}
// Disk::hit (disk.rs:31-62)
__device__ __forceinline__ bool disk_test(const DevInstance &I, float3 o, float3 d, float t0, float t1, float tmax, float &t_out) {
  float radius = I.origin_size0.w;
  float3 tmp_o = o - f3(I.origin_size0);
  if (d.z == 0.0f) return false;
  float t = (-tmp_o.z) / d.z;
  if (t <= t0 || t > t1 || t >= tmax) return false;
  float xh = __fadd_rn(tmp_o.x, __fmul_rn(t, d.x)), yh = __fadd_rn(tmp_o.y, __fmul_rn(t, d.y));
  if (__fadd_rn(__fmul_rn(xh, xh), __fmul_rn(yh, yh)) > __fmul_rn(radius, radius)) return false;
  t_out = t;
  return true;
}
__device__ __forceinline__ void disk_finish(const DevInstance &I, float3 o, float3 d, float t, LocalHit &h) {
  float3 n = f3(0, 0, 1);
  if (dot(d, n) > 0.0f && (I.flags & DI_TWO_SIDED)) n = -n;
  h.t = t;
  h.p = o + d * t;
  h.u = h.v = 0.0f;
  h.n = n;
  h.material = RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
}

__device__ __forceinline__ float3 tri_shuffle(float3 v, uint32_t kz) {  // mesh.rs:12-19
  return kz == 0 ? f3(v.y, v.z, v.x) : (kz == 1 ? f3(v.z, v.x, v.y) : v);
}
// MeshTriangleRef::hit (mesh.rs:67-198): watertight test, f64 fallback on zero edge functions.
// Outputs t and the three barycentrics exactly as the reference computes them.
__device__ __forceinline__ bool tri_test(float3 p0, float3 p1, float3 p2, float3 o, float3 d, float t0, float t1, float &t_out,
                                         float &b0, float &b1, float &b2) {
  float3 p0t = p0 - o, p1t = p1 - o, p2t = p2 - o;
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  float mx = fmaxf(fmaxf(ax, ay), az);
  uint32_t kz = 0;
  if (ay >= mx) kz = 1;
  if (az >= mx) kz = 2;  // ties -> highest axis index (mesh.rs:80-85)
  float3 dd = tri_shuffle(d, kz);
  p0t = tri_shuffle(p0t, kz);
  p1t = tri_shuffle(p1t, kz);
  p2t = tri_shuffle(p2t, kz);
  float sx = -dd.x / dd.z, sy = -dd.y / dd.z, sz = 1.0f / dd.z;
  p0t.x = __fadd_rn(p0t.x, __fmul_rn(sx, p0t.z));
  p1t.x = __fadd_rn(p1t.x, __fmul_rn(sx, p1t.z));
  p2t.x = __fadd_rn(p2t.x, __fmul_rn(sx, p2t.z));
  p0t.y = __fadd_rn(p0t.y, __fmul_rn(sy, p0t.z));
  p1t.y = __fadd_rn(p1t.y, __fmul_rn(sy, p1t.z));
  p2t.y = __fadd_rn(p2t.y, __fmul_rn(sy, p2t.z));
  float e0 = __fsub_rn(__fmul_rn(p1t.x, p2t.y), __fmul_rn(p1t.y, p2t.x));
  float e1 = __fsub_rn(__fmul_rn(p2t.x, p0t.y), __fmul_rn(p2t.y, p0t.x));
  float e2 = __fsub_rn(__fmul_rn(p0t.x, p1t.y), __fmul_rn(p0t.y, p1t.x));
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    e0 = (float)__dsub_rn(__dmul_rn((double)p2t.y, (double)p1t.x), __dmul_rn((double)p2t.x, (double)p1t.y));
    e1 = (float)__dsub_rn(__dmul_rn((double)p0t.y, (double)p2t.x), __dmul_rn((double)p0t.x, (double)p2t.y));
    e2 = (float)__dsub_rn(__dmul_rn((double)p1t.y, (double)p0t.x), __dmul_rn((double)p1t.x, (double)p0t.y));
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = __fadd_rn(__fadd_rn(e0, e1), e2);
  if (det == 0.0f) return false;
  float z0 = __fmul_rn(p0t.z, sz), z1 = __fmul_rn(p1t.z, sz), z2 = __fmul_rn(p2t.z, sz);
  float t_scaled = __fadd_rn(__fadd_rn(__fmul_rn(e0, z0), __fmul_rn(e1, z1)), __fmul_rn(e2, z2));
  float lo = __fmul_rn(t0, det), hi = __fmul_rn(t1, det);
  if ((det < 0.0f && (t_scaled >= lo || t_scaled < hi)) || (det > 0.0f && (t_scaled <= lo || t_scaled > hi))) return false;
  float inv_det = 1.0f / det;
  b0 = __fmul_rn(e0, inv_det);
  b1 = __fmul_rn(e1, inv_det);
  b2 = __fmul_rn(e2, inv_det);
  t_out = __fmul_rn(t_scaled, inv_det);
  return true;
}

// Robust slab test against one child box; tnear returned for ordering. Conservative (never rejects a
// box the exact test accepts): the far plane is widened by 2*gamma(3) as in PBRT's Bounds3::IntersectP. This is
// the one place that uses explicit FMAs (t = b * inv_d - o * inv_d): it only prunes, it never decides a hit, so
// its rounding is parity-irrelevant.
// Zero direction components: the reference leaves such an axis unconstrained (aabb.rs:41-45: d == 0 -> (0, +inf)).
// b * (+-inf) - o * (+-inf) does NOT do that by itself (it is NaN only when b and o have the same sign, +-inf
// otherwise, and fmin/fmax then keep the infinite plane and reject the box), so slab_recip() poisons o * inv_d with
// a NaN on those axes: both plane distances become NaN and fminf / fmaxf drop the axis, at no cost per node.
__device__ __forceinline__ void slab_recip(float3 o, float3 d, float3 &inv, float3 &oinv);
__device__ __forceinline__ bool slab_test(float3 bmin, float3 bmax, float3 oinv, float3 inv_d, float tmax, float &tnear) {
  float tx0 = fmaf(bmin.x, inv_d.x, -oinv.x), tx1 = fmaf(bmax.x, inv_d.x, -oinv.x);
  float ty0 = fmaf(bmin.y, inv_d.y, -oinv.y), ty1 = fmaf(bmax.y, inv_d.y, -oinv.y);
  float tz0 = fmaf(bmin.z, inv_d.z, -oinv.z), tz1 = fmaf(bmax.z, inv_d.z, -oinv.z);
  float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
  float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
  tf *= 1.000001f;  // >= 1 + 2*gamma(3), plus headroom for the ~1 ulp reciprocal
  tnear = tn;
  return tn <= tf;
}

// Closest-hit result of a BVH query.
struct TraceHit {
  float t;
  uint32_t inst, prim;
};

// Tie-break keys reproduce the reference's candidate-loop semantics (accelerator/mod.rs:143-175,
// mesh.rs:331-357): candidates are visited in flat-BVH order and replace the current best when
// `t <= closest` (rect, disk, triangle) or `t < closest` (sphere). Among equal-t candidates the winner
// is therefore the LAST non-strict one in order, else the FIRST strict one. key = larger wins.
__device__ __forceinline__ uint64_t tie_key(bool strict, uint32_t inst_order, uint32_t tri_order) {
  uint64_t k = strict ? (uint64_t)(0x7FFFFFFFu - inst_order) : (0x80000000ull | inst_order);
  return (k << 32) | tri_order;
}

#define RPT_STACK_SIZE 64

// Two-level closest-hit traversal (replaces World::hit -> Accelerator::hit -> FlatBVH::traverse ->
// Instance::hit -> Mesh::hit; world/mod.rs:166, accelerator/mod.rs:86-178, lbvh.rs:172-213,
// instance.rs:75-133, mesh.rs:314-360). Unlike the reference (F8) it prunes by the closest hit so far;
// results are identical because pruned boxes cannot contain a closer hit.
// `stack` is this thread's slice of the CTA's shared-memory traversal stack, strided by `stride`.
// Work counters for the roofline accounting (bytes fetched per ray = nodes * 64 B + triangles * 48 B + instances * 144 B).
struct TraceWork {
  uint32_t nodes, tris, insts;
};

#define RPT_DONE ((int)0x80000000)

// 1/x for the slab test only (MUFU.RCP, ~1 ulp): it orders and prunes, it never decides a hit.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void slab_recip(float3 o, float3 d, float3 &inv, float3 &oinv) {
  const float qnan = __int_as_float(0x7fffffff);
  inv = f3(rcp_approx(d.x), rcp_approx(d.y), rcp_approx(d.z));
  oinv = f3(d.x == 0.0f ? qnan : o.x * inv.x, d.y == 0.0f ? qnan : o.y * inv.y, d.z == 0.0f ? qnan : o.z * inv.z);
}

// Watertight-test constants that depend only on the ray (mesh.rs:76-100 computes them per triangle; hoisting
// them is exact: same inputs, same roundings).
struct TriRay {
  uint32_t kz;
  float sx, sy, sz;
};
__device__ __forceinline__ TriRay tri_ray_setup(float3 d) {
  TriRay r;
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  float mx = fmaxf(fmaxf(ax, ay), az);
  r.kz = 0;
  if (ay >= mx) r.kz = 1;
  if (az >= mx) r.kz = 2;  // ties -> highest axis index (mesh.rs:80-85)
  float3 dd = tri_shuffle(d, r.kz);
  r.sx = -dd.x / dd.z;
  r.sy = -dd.y / dd.z;
  r.sz = 1.0f / dd.z;
  return r;
}
// MeshTriangleRef::hit with the per-ray part precomputed; identical arithmetic to tri_test().
__device__ __forceinline__ bool tri_test_pre(float3 p0, float3 p1, float3 p2, float3 o, const TriRay &tr, float t0, float t1, float &t_out,
                                             float &b0, float &b1, float &b2) {
  float3 p0t = tri_shuffle(p0 - o, tr.kz), p1t = tri_shuffle(p1 - o, tr.kz), p2t = tri_shuffle(p2 - o, tr.kz);
  p0t.x = __fadd_rn(p0t.x, __fmul_rn(tr.sx, p0t.z));
  p1t.x = __fadd_rn(p1t.x, __fmul_rn(tr.sx, p1t.z));
  p2t.x = __fadd_rn(p2t.x, __fmul_rn(tr.sx, p2t.z));
  p0t.y = __fadd_rn(p0t.y, __fmul_rn(tr.sy, p0t.z));
  p1t.y = __fadd_rn(p1t.y, __fmul_rn(tr.sy, p1t.z));
  p2t.y = __fadd_rn(p2t.y, __fmul_rn(tr.sy, p2t.z));
  float e0 = __fsub_rn(__fmul_rn(p1t.x, p2t.y), __fmul_rn(p1t.y, p2t.x));
  float e1 = __fsub_rn(__fmul_rn(p2t.x, p0t.y), __fmul_rn(p2t.y, p0t.x));
  float e2 = __fsub_rn(__fmul_rn(p0t.x, p1t.y), __fmul_rn(p0t.y, p1t.x));
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    e0 = (float)__dsub_rn(__dmul_rn((double)p2t.y, (double)p1t.x), __dmul_rn((double)p2t.x, (double)p1t.y));
    e1 = (float)__dsub_rn(__dmul_rn((double)p0t.y, (double)p2t.x), __dmul_rn((double)p0t.x, (double)p2t.y));
    e2 = (float)__dsub_rn(__dmul_rn((double)p1t.y, (double)p0t.x), __dmul_rn((double)p1t.x, (double)p0t.y));
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = __fadd_rn(__fadd_rn(e0, e1), e2);
  if (det == 0.0f) return false;
  float z0 = __fmul_rn(p0t.z, tr.sz), z1 = __fmul_rn(p1t.z, tr.sz), z2 = __fmul_rn(p2t.z, tr.sz);
  float t_scaled = __fadd_rn(__fadd_rn(__fmul_rn(e0, z0), __fmul_rn(e1, z1)), __fmul_rn(e2, z2));
  float lo = __fmul_rn(t0, det), hi = __fmul_rn(t1, det);
  if ((det < 0.0f && (t_scaled >= lo || t_scaled < hi)) || (det > 0.0f && (t_scaled <= lo || t_scaled > hi))) return false;
  float inv_det = 1.0f / det;
  b0 = __fmul_rn(e0, inv_det);
  b1 = __fmul_rn(e1, inv_det);
  b2 = __fmul_rn(e2, inv_det);
  t_out = __fmul_rn(t_scaled, inv_det);
  return true;
}

// Two-level closest-hit traversal (replaces World::hit -> Accelerator::hit -> FlatBVH::traverse ->
// Instance::hit -> Mesh::hit; world/mod.rs:166, accelerator/mod.rs:86-178, lbvh.rs:172-213,
// instance.rs:75-133, mesh.rs:314-360). Unlike the reference (F8) it prunes by the closest hit so far;
// results are identical because pruned boxes cannot contain a closer hit. State lives in registers.
//
// WIDE = true walks the four-wide collapse of the same trees (DevScene::nodes4, built by rpt::collapse_bvh4). A node is 8 float4:
// rows 0-2 the four children's box minima (x, y, z), rows 3-5 their maxima, row 6 the four child refs (RPT_DONE = empty slot),
// row 7 padding to one 128-byte line. The near / far plane of every child along an axis is the min or the max row depending
// only on the sign of the ray direction, so the six plane rows are fetched through three per-ray row offsets and the slab test
// needs no min / max per box; the (up to four) children that are hit are visited nearest first. Boxes, leaves and the
// accept() rule are those of the two-wide tree, so the result is the same hit (tests/test_gpu_parity.py::test_bvh4_*).
// TWO_LEVEL = false is the walk for scenes whose TLAS leaves are all triangles and analytic shapes (no transformed mesh
// instance: Cornell, the furnace, the HDR scenes): the instance-local ray, the BLAS bookkeeping and the change-of-space code
// are compiled out, which frees ~10 registers of state.
template <bool WIDE, bool TWO_LEVEL = true>
struct TravT {
  float3 o, d;       // world-space ray
  float3 ro, rd;     // current-space ray (instance-local inside a BLAS)
  float3 inv, oinv;  // slab-test reciprocals of the current-space ray
  TriRay tr;         // watertight-test constants of the current-space ray
  float tmax, closest;
  uint64_t best_key;
  int cur, sp, blas_base;
  uint32_t cur_inst, cur_inst_order, cur_tri_base;
  uint32_t near_rows;  // WIDE: row of the near plane per axis, 2 bits each at bits 0 / 8 / 16 (x: 0 or 3, y: 1 or 4, z: 2 or 5)
  bool found;
  TraceHit out;

  // The watertight-test constants (three IEEE divisions) are computed eagerly only for the world-space ray at init(), where the
  // whole warp does it together. After a change of space (entering or leaving a mesh instance: 1.5 of each per ray on the
  // instanced 10 M-triangle scene, taken by 2.5-4.3 lanes at a time) they are only marked stale (kz = 3) and rebuilt by the
  // first triangle test that needs them - many instance visits end in the BLAS's boxes without reaching a triangle.
  template <bool EAGER>
  __device__ __forceinline__ void set_space(float3 no, float3 nd) {
    ro = no;
    rd = nd;
    slab_recip(ro, rd, inv, oinv);
#ifdef RPT_EAGER_TRI_SETUP  // (A/B build: the constants rebuilt at every change of space)
    tr = tri_ray_setup(rd);
#else
    if (EAGER) tr = tri_ray_setup(rd);
    else tr.kz = 3u;
#endif
    if (WIDE) near_rows = (inv.x < 0.0f ? 3u : 0u) | (inv.y < 0.0f ? 4u : 1u) << 8 | (inv.z < 0.0f ? 5u : 2u) << 16;
  }
  __device__ __forceinline__ void init(const DevScene &S, float3 o_, float3 d_, float tmax_) {
    o = o_;
    d = d_;
    tmax = closest = tmax_;
    best_key = 0;
    found = false;
    out.t = RPT_INF;
    out.inst = RPT_NONE;
    out.prim = RPT_NONE;
    sp = 0;
    blas_base = 0;
    cur = WIDE ? S.tlas_root4 : S.tlas_root;
    cur_inst = RPT_NONE;
    cur_inst_order = cur_tri_base = 0;
    set_space<true>(o, d);
  }
  // Pops the next node ref; leaving a BLAS (its stack segment is exhausted) restores the world-space ray.
  __device__ __forceinline__ int pop(const int *stack, int stride) {
    if (TWO_LEVEL && cur_inst != RPT_NONE && sp == blas_base) {
      set_space<false>(o, d);
      cur_inst = RPT_NONE;
    }
    if (sp == 0) return RPT_DONE;
    --sp;
    return stack[sp * stride];
  }
  __device__ __forceinline__ bool accept(float t, uint64_t key, uint32_t inst, uint32_t prim) {
    if (!found || t < closest || key > best_key) {
      closest = t;
      best_key = key;
      found = true;
      out.t = t;
      out.inst = inst;
      out.prim = prim;
      return true;
    }
    return false;
  }

  // "while-while": the inner loop keeps every lane of the warp on the node-test code until each has reached a
  // leaf (or run out of work); leaves are then processed together, so the warp reconverges at the end of the
  // inner loop instead of drifting apart iteration by iteration.
  template <bool ANY_HIT, bool STATS>
  __device__ __forceinline__ void run(const DevScene &S, int *stack, int stride, TraceWork &work) {
    while (!step<ANY_HIT, STATS>(S, stack, stride, work)) {
    }
  }
  // One iteration of the outer loop: descend to the next leaf, process it, fetch the next node ref.
  // Returns true when the ray is finished (nothing left to visit, or ANY_HIT accepted a hit).
  template <bool ANY_HIT, bool STATS>
  __device__ __forceinline__ bool step(const DevScene &S, int *stack, int stride, TraceWork &work) {
    {
      // ---- phase 1: descend through inner nodes (inner refs are >= 0)
      while (WIDE && cur >= 0) {
        RPT_STAT(work.nodes++);
        const float4 *np = S.nodes4 + 8 * (size_t)cur;
        const uint32_t rx = near_rows & 0xFFu, ry = (near_rows >> 8) & 0xFFu, rz = near_rows >> 16;
        const float4 nx = __ldg(np + rx), ny = __ldg(np + ry), nz = __ldg(np + rz);
        const float4 fx = __ldg(np + (3u - rx)), fy = __ldg(np + (5u - ry)), fz = __ldg(np + (7u - rz));
        const int4 ch = __ldg(reinterpret_cast<const int4 *>(np + 6));
        // same planes, same fmaf as slab_test(); a zero direction component turns both of its planes into NaN, which fmaxf / fminf drop
        const float lim = closest;
#define RPT_WIDE_CHILD(c, T, R)                                                                                                          \
  {                                                                                                                                       \
    const float tn = fmaxf(fmaxf(fmaf(nx.c, inv.x, -oinv.x), fmaf(ny.c, inv.y, -oinv.y)), fmaxf(fmaf(nz.c, inv.z, -oinv.z), 0.0f));       \
    const float tf = fminf(fminf(fmaf(fx.c, inv.x, -oinv.x), fmaf(fy.c, inv.y, -oinv.y)), fminf(fmaf(fz.c, inv.z, -oinv.z), lim)) * 1.000001f; \
    T = (tn <= tf && ch.c != RPT_DONE) ? fminf(tn, 3.402823466e38f) : RPT_INF;                                                                                   \
    R = ch.c;                                                                                                                             \
  }
        float t0, t1, t2, t3;
        int r0, r1, r2, r3;
        RPT_WIDE_CHILD(x, t0, r0)
        RPT_WIDE_CHILD(y, t1, r1)
        RPT_WIDE_CHILD(z, t2, r2)
        RPT_WIDE_CHILD(w, t3, r3)
#undef RPT_WIDE_CHILD
        // sort the four (t, ref) pairs by t: misses (t = +inf; a hit's t is clamped to FLT_MAX) end up last
#define RPT_WIDE_CSWAP(ta, ra, tb, rb)    \
  {                                       \
    const bool sw = tb < ta;              \
    const float tlo = sw ? tb : ta;       \
    const int rlo = sw ? rb : ra;         \
    tb = sw ? ta : tb;                    \
    rb = sw ? ra : rb;                    \
    ta = tlo;                             \
    ra = rlo;                             \
  }
        RPT_WIDE_CSWAP(t0, r0, t1, r1)
        RPT_WIDE_CSWAP(t2, r2, t3, r3)
        RPT_WIDE_CSWAP(t0, r0, t2, r2)
        RPT_WIDE_CSWAP(t1, r1, t3, r3)
        RPT_WIDE_CSWAP(t1, r1, t2, r2)
#undef RPT_WIDE_CSWAP
        if (t0 == RPT_INF) {
          cur = pop(stack, stride);
        } else {
          if (t3 != RPT_INF) stack[(sp++) * stride] = r3;
          if (t2 != RPT_INF) stack[(sp++) * stride] = r2;
          if (t1 != RPT_INF) stack[(sp++) * stride] = r1;
          cur = r0;
        }
      }
      while (!WIDE && cur >= 0) {
        RPT_STAT(work.nodes++);
        const float4 *np = reinterpret_cast<const float4 *>(S.nodes + cur);
        float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
        int4 ch = __ldg(reinterpret_cast<const int4 *>(np + 3));
        float tl, trr;
        bool hl = slab_test(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), oinv, inv, closest, tl);
        bool hr = slab_test(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), oinv, inv, closest, trr);
        if (hl && hr) {
          bool left_first = tl <= trr;
          stack[sp * stride] = left_first ? ch.y : ch.x;
          ++sp;
          cur = left_first ? ch.x : ch.y;
        } else if (hl || hr) {
          cur = hl ? ch.x : ch.y;
        } else {
          cur = pop(stack, stride);
        }
      }
      if (cur == RPT_DONE) return true;

      // ---- phase 2: a leaf
      uint32_t idx = (uint32_t)(~cur);
      bool is_tri = true;
      uint32_t tri = 0, tri_local = idx, hit_inst = cur_inst, inst_order = cur_inst_order;
      int next = 0;
      bool have_next = false;
      if (TWO_LEVEL && cur_inst != RPT_NONE) {
        tri = cur_tri_base + idx;  // BLAS leaf: a triangle of the current mesh instance
      } else {
        uint4 lf = __ldg(S.tlas_leaves + idx);
        hit_inst = lf.x;
        inst_order = lf.w;
        if (lf.y != RPT_NONE) {
          tri_local = lf.y;  // flattened triangle of an untransformed mesh instance: world ray as is
          tri = lf.z;
        } else {
          // TLAS leaf: a whole instance
          is_tri = false;
          RPT_STAT(work.insts++);
          const DevInstance &I = S.instances[hit_inst];
          uint32_t flags = I.flags;
          float3 lo = o, ld = d;
          if (flags & DI_HAS_TRANSFORM) {  // instance.rs:89-95: direction is NOT renormalised, t is shared
            lo = xform_point(I.rev, o);
            ld = xform_vec(I.rev, d);
          }
          uint32_t kind = flags & DI_KIND_MASK;
          if (TWO_LEVEL && kind == RPT_AGG_MESH) {  // (a scene that gets the one-level walk has no such leaf)
            blas_base = sp;
            set_space<false>(lo, ld);
            cur_inst = hit_inst;
            cur_inst_order = inst_order;
            cur_tri_base = I.tri_base;
            next = WIDE ? I.blas_root4 : I.blas_root;
            have_next = true;
          } else {
            float t;
            bool hit;
            // closest-so-far is passed as t1 exactly as the reference's candidate loop does; a candidate
            // that passes with t == closest is resolved by tie_key (a strict sphere never gets that far).
            if (kind == RPT_AGG_RECT)
              hit = rect_test(I, lo, ld, 0.0f, closest, tmax, t);
            else if (kind == RPT_AGG_SPHERE)
              hit = sphere_test(I, lo, ld, 0.0f, closest, tmax, t);
            else
              hit = disk_test(I, lo, ld, 0.0f, closest, tmax, t);
            if (hit && accept(t, tie_key(kind == RPT_AGG_SPHERE, inst_order, 0), hit_inst, 0) && ANY_HIT) return true;
          }
        }
      }
      if (is_tri) {
        RPT_STAT(work.tris++);
        const float4 *tv = S.tri_verts + 3 * (size_t)tri;
        float4 v0 = __ldg(tv), v1 = __ldg(tv + 1), v2 = __ldg(tv + 2);
        float t, b0, b1, b2;
#ifndef RPT_EAGER_TRI_SETUP
        if (tr.kz > 2u) tr = tri_ray_setup(rd);
#endif
        if (tri_test_pre(f3(v0), f3(v1), f3(v2), ro, tr, 0.0f, closest, t, b0, b1, b2) &&
            accept(t, tie_key(false, inst_order, __float_as_uint(v1.w)), hit_inst, tri_local) && ANY_HIT)
          return true;
      }
      cur = have_next ? next : pop(stack, stride);
      return cur == RPT_DONE;
    }
  }
};

using Trav = TravT<false>;

template <bool ANY_HIT, bool STATS, bool WIDE = false, bool TWO_LEVEL = true>
__device__ __forceinline__ bool trace_ray(const DevScene &S, float3 o, float3 d, float tmax, int *stack, int stride, TraceHit &out,
                                          TraceWork &work) {
  TravT<WIDE, TWO_LEVEL> t;
  t.init(S, o, d, tmax);
  t.template run<ANY_HIT, STATS>(S, stack, stride, work);
  out = t.out;
  return t.found;
}

// MeshTriangleRef::hit on vertices that are ALREADY translated by the ray origin and permuted (tri_shuffle commutes with
// the subtraction component by component, so the arithmetic below is the arithmetic of tri_test_pre, rounding for rounding).
__device__ __forceinline__ bool tri_test_shuffled(float3 p0t, float3 p1t, float3 p2t, const TriRay &tr, float t0, float t1, float &t_out) {
  p0t.x = __fadd_rn(p0t.x, __fmul_rn(tr.sx, p0t.z));
  p1t.x = __fadd_rn(p1t.x, __fmul_rn(tr.sx, p1t.z));
  p2t.x = __fadd_rn(p2t.x, __fmul_rn(tr.sx, p2t.z));
  p0t.y = __fadd_rn(p0t.y, __fmul_rn(tr.sy, p0t.z));
  p1t.y = __fadd_rn(p1t.y, __fmul_rn(tr.sy, p1t.z));
  p2t.y = __fadd_rn(p2t.y, __fmul_rn(tr.sy, p2t.z));
  float e0 = __fsub_rn(__fmul_rn(p1t.x, p2t.y), __fmul_rn(p1t.y, p2t.x));
  float e1 = __fsub_rn(__fmul_rn(p2t.x, p0t.y), __fmul_rn(p2t.y, p0t.x));
  float e2 = __fsub_rn(__fmul_rn(p0t.x, p1t.y), __fmul_rn(p0t.y, p1t.x));
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    e0 = (float)__dsub_rn(__dmul_rn((double)p2t.y, (double)p1t.x), __dmul_rn((double)p2t.x, (double)p1t.y));
    e1 = (float)__dsub_rn(__dmul_rn((double)p0t.y, (double)p2t.x), __dmul_rn((double)p0t.x, (double)p2t.y));
    e2 = (float)__dsub_rn(__dmul_rn((double)p1t.y, (double)p0t.x), __dmul_rn((double)p1t.x, (double)p0t.y));
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = __fadd_rn(__fadd_rn(e0, e1), e2);
  if (det == 0.0f) return false;
  float z0 = __fmul_rn(p0t.z, tr.sz), z1 = __fmul_rn(p1t.z, tr.sz), z2 = __fmul_rn(p2t.z, tr.sz);
  float t_scaled = __fadd_rn(__fadd_rn(__fmul_rn(e0, z0), __fmul_rn(e1, z1)), __fmul_rn(e2, z2));
  float lo = __fmul_rn(t0, det), hi = __fmul_rn(t1, det);
  if ((det < 0.0f && (t_scaled >= lo || t_scaled < hi)) || (det > 0.0f && (t_scaled <= lo || t_scaled > hi))) return false;
  t_out = __fmul_rn(t_scaled, 1.0f / det);
  return true;
}

// Small-scene closest / any hit: no BVH. Scenes of a few dozen primitives (the Cornell box: 30 triangles and a rect)
// fit shared memory whole, and a BVH walk over them runs 11-15 of 32 lanes (ncu, profiles/r01_final_ncu_kernels.csv)
// because every lane is at a different node. Here the warp walks the leaf list in lockstep - leaf k is the same
// primitive for all 32 lanes, fetched as a shared-memory broadcast - so no lane idles, there is no stack, and the
// result is the same by construction: the accept rule (closest t, then the reference's candidate-order tie key) does not
// depend on the order candidates are visited in. A lane reads its triangle already permuted for its own dominant ray
// axis (three copies per triangle, 144 B), which removes the per-vertex selects of tri_shuffle from the inner loop.
// All 32 lanes of the warp must call run() (ANY_HIT votes); `active` masks the lanes that carry a ray.
#define RPT_SMALL_MAX 64u
struct SmallTrav {
  float closest;
  uint64_t best_key;
  bool found;
  TraceHit out;
  __device__ __forceinline__ void init(float tmax) {
    closest = tmax;
    best_key = 0;
    found = false;
    out.t = RPT_INF;
    out.inst = RPT_NONE;
    out.prim = RPT_NONE;
  }
  __device__ __forceinline__ bool accept(float t, uint64_t key, uint32_t inst, uint32_t prim) {
    if (!found || t < closest || key > best_key) {
      closest = t;
      best_key = key;
      found = true;
      out.t = t;
      out.inst = inst;
      out.prim = prim;
      return true;
    }
    return false;
  }
  // s_tris: the CTA's shared-memory copy of DevScene::small_tris. Returns with `found` / `out` / `closest` / `best_key` set.
  template <bool ANY_HIT, bool STATS>
  __device__ __forceinline__ void run(const DevScene &S, const float4 *s_tris, float3 o, float3 d, float tmax, bool active, TraceWork &work) {
    const TriRay tr = tri_ray_setup(d);
    const float3 os = tri_shuffle(o, tr.kz);
    bool alive = active;
    const uint32_t ntri = S.small_ntri, n = S.small_n;
    for (uint32_t k = 0; k < ntri; ++k) {
      if (ANY_HIT && !__any_sync(0xFFFFFFFFu, alive)) return;
      if (alive) {
        RPT_STAT(work.tris++);
        const float4 *tv = s_tris + 9u * k + 3u * tr.kz;
        float4 a = tv[0], b = tv[1], c = tv[2];
        float t;
        if (tri_test_shuffled(f3(a) - os, f3(b) - os, f3(c) - os, tr, 0.0f, closest, t)) {
          uint4 lf = __ldg(S.small_leaves + k);
          if (accept(t, tie_key(false, lf.w, lf.z), lf.x, lf.y) && ANY_HIT) alive = false;
        }
      }
    }
    for (uint32_t k = ntri; k < n; ++k) {
      if (ANY_HIT && !__any_sync(0xFFFFFFFFu, alive)) return;
      if (alive) {
        RPT_STAT(work.insts++);
        uint4 lf = __ldg(S.small_leaves + k);
        const DevInstance &I = S.instances[lf.x];
        uint32_t flags = I.flags;
        float3 lo = o, ld = d;
        if (flags & DI_HAS_TRANSFORM) {  // instance.rs:89-95
          lo = xform_point(I.rev, o);
          ld = xform_vec(I.rev, d);
        }
        uint32_t kind = flags & DI_KIND_MASK;
        float t;
        bool hit;
        if (kind == RPT_AGG_RECT) {
          hit = rect_test(I, lo, ld, 0.0f, closest, tmax, t);
        } else if (kind == RPT_AGG_SPHERE) {
          hit = sphere_test(I, lo, ld, 0.0f, closest, tmax, t);
        } else {
          // A disk's bounding box is SMALLER than the disk (disk.rs:23-28: half extent radius / 2): the part outside it
          // is unreachable through the reference's BVH, so the disk is gated on that box here as a BVH leaf would be.
          const float *bx = S.small_boxes + 6u * (k - ntri);
          float3 winv, woinv;
          slab_recip(o, d, winv, woinv);
          float tn;
          hit = slab_test(f3(__ldg(bx), __ldg(bx + 1), __ldg(bx + 2)), f3(__ldg(bx + 3), __ldg(bx + 4), __ldg(bx + 5)), woinv, winv, closest, tn) &&
                disk_test(I, lo, ld, 0.0f, closest, tmax, t);
        }
        if (hit && accept(t, tie_key(kind == RPT_AGG_SPHERE, lf.w, 0), lf.x, 0) && ANY_HIT) alive = false;
      }
    }
  }
};

// Full hit record of a known (instance, primitive, t): Instance::hit's output (instance.rs:96-116).
struct SurfaceHit {
  float3 p, n;
  float u, v;
  uint32_t material;
};
__device__ __forceinline__ void reconstruct_hit(const DevScene &S, float3 o, float3 d, const TraceHit &th, SurfaceHit &sh) {
  const DevInstance &I = S.instances[th.inst];
  uint32_t flags = I.flags;
  float3 lo = o, ld = d;
  if (flags & DI_HAS_TRANSFORM) {
    lo = xform_point(I.rev, o);
    ld = xform_vec(I.rev, d);
  }
  LocalHit h;
  uint32_t kind = flags & DI_KIND_MASK;
  if (kind == RPT_AGG_MESH) {
    uint32_t tri = I.tri_base + th.prim;
    const float4 *tv = S.tri_verts + 3 * (size_t)tri;
    float4 v0 = __ldg(tv), v1 = __ldg(tv + 1), v2 = __ldg(tv + 2);
    float3 p0 = f3(v0), p1 = f3(v1), p2 = f3(v2);
    float t, b0, b1, b2;
    // same arithmetic as the trace kernel -> same t and barycentrics (deterministic); t1 = +inf
    tri_test(p0, p1, p2, lo, ld, 0.0f, RPT_INF, t, b0, b1, b2);
    float3 n;
    if (I.has_normals) {
      const float4 *tn = S.tri_normals + 3 * (size_t)tri;
      n = b0 * f3(__ldg(tn)) + b1 * f3(__ldg(tn + 1)) + b2 * f3(__ldg(tn + 2));  // mesh.rs:169-179
    } else {
      n = normalized(cross(p0 - p2, p1 - p2));  // mesh.rs:164-166
    }
    h.t = th.t;
    h.p = b0 * p0 + b1 * p1 + b2 * p2;
    h.u = h.v = 0.0f;
    h.n = normalized(n);  // HitRecord::new (hittable.rs:34)
    h.material = __float_as_uint(v0.w);
  } else if (kind == RPT_AGG_RECT) {
    rect_finish(I, lo, ld, th.t, h);
  } else if (kind == RPT_AGG_SPHERE) {
    sphere_finish(I, lo, ld, th.t, h);
  } else {
    disk_finish(I, lo, ld, th.t, h);
  }
  if (flags & DI_HAS_TRANSFORM) {
    h.n = normalized(xform_vec_transposed(I.rev, h.n));
    h.p = xform_point(I.fwd, h.p);
  }
  sh.p = h.p;
  sh.n = h.n;
  sh.u = h.u;
  sh.v = h.v;
  sh.material = I.material != RPT_NONE ? I.material : h.material;
}

// ---------------------------------------------------------------------------------------------
// light sampling (Hittable::sample / psa_pdf through Instance)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float finite_or_zero(float p) { return isfinite(p) ? p : 0.0f; }
// Instance::sample (instance.rs:134-141) over AARect / Sphere / Disk (rect.rs:113-155, sphere.rs:88-131, disk.rs:63-91)
__device__ __forceinline__ void instance_sample(const DevInstance &I, float sx, float sy, float3 from, float3 &dir, float &pdf) {
  uint32_t flags = I.flags;
  uint32_t kind = flags & DI_KIND_MASK;
  float3 f = (flags & DI_HAS_TRANSFORM) ? xform_point(I.rev, from) : from;
  float3 org = f3(I.origin_size0);
  float3 direction;
  float p;
  if (kind == RPT_AGG_SPHERE) {
    float radius = I.origin_size0.w;
    float3 n = random_on_unit_sphere(sx, sy);
    float3 point = org + radius * n;
    float area_pdf = 1.0f / (radius * radius * 4.0f * RPT_PI);
    direction = point - f;
    float ndd = fabsf(dot(n, normalized(direction)));
    p = area_pdf * norm_squared(direction) / ndd;
  } else {
    uint32_t axis = (flags >> DI_AXIS_SHIFT) & 3u;
    float3 n = kind == RPT_AGG_RECT ? axis_vec(axis) : f3(0, 0, 1);
    float x = sx;
    if (flags & DI_TWO_SIDED) {
      float resc;
      bool first = choose(x, 0.5f, resc);
      x = resc;
      n = n * (first ? -1.0f : 1.0f);
    }
    float3 point;
    float area;
    if (kind == RPT_AGG_RECT) {
      point = org + shuffle_axis(f3((x - 0.5f) * I.origin_size0.w, (sy - 0.5f) * I.size1, 0.0f), axis);
      area = I.origin_size0.w * I.size1;
    } else {
      float radius = I.origin_size0.w;
      point = org + radius * random_in_unit_disk(x, sy);
      area = RPT_PI * radius * radius;
    }
    direction = point - f;
    float cos_i = dot(n, normalized(direction));
    p = (1.0f / area) * norm_squared(direction) / fabsf(cos_i);  // Area -> SolidAngle
  }
  dir = normalized(direction);
  pdf = finite_or_zero(p);
  if (flags & DI_HAS_TRANSFORM) dir = normalized(xform_vec(I.fwd, dir));
}
// Instance::psa_pdf (instance.rs:154-170; rect.rs:156-173, sphere.rs:132-152, disk.rs:92-104)
__device__ __forceinline__ float instance_psa_pdf(const DevInstance &I, float cos_o, float cos_i, float3 from, float3 to) {
  uint32_t flags = I.flags;
  if (flags & DI_HAS_TRANSFORM) {  // to_world (reference quirk Q11)
    from = xform_point(I.fwd, from);
    to = xform_point(I.fwd, to);
  }
  float d2 = norm_squared(to - from);
  uint32_t kind = flags & DI_KIND_MASK;
  if (kind == RPT_AGG_RECT) {
    float area_pdf = 1.0f / (I.origin_size0.w * I.size1);
    return (area_pdf * d2 / fabsf(cos_i)) / fabsf(cos_o);
  } else if (kind == RPT_AGG_SPHERE) {
    float area_pdf = 1.0f / (I.origin_size0.w * I.origin_size0.w * 4.0f * RPT_PI);
    return area_pdf * d2 / fabsf(cos_i * cos_o);
  } else if (kind == RPT_AGG_DISK) {
    float area = RPT_PI * I.origin_size0.w * I.origin_size0.w;
    return d2 / ((fabsf(cos_o) * fabsf(cos_i) + 0.00001f) * area);
  }
  return 0.0f;
}

// ---------------------------------------------------------------------------------------------
// materials (materials/{lambertian,ggx,diffuse_light,sharp_light}.rs)
// ---------------------------------------------------------------------------------------------
struct Bsdf {
  float f, pdf;
};
__device__ __forceinline__ float3 ggx_reflect(float3 wi, float3 n) {  // ggx.rs:3-6
  float3 w = -wi;
  return normalized(w - 2.0f * dot(w, n) * n);
}
__device__ __forceinline__ bool ggx_refract(float3 wi, float3 n, float eta, float3 &out) {  // ggx.rs:8-17
  float cos_i = dot(wi, n);
  float sin2i = fmaxf(1.0f - cos_i * cos_i, 0.0f);
  float sin2t = eta * eta * sin2i;
  if (sin2t >= 1.0f) return false;
  float cos_t = sqrtf(1.0f - sin2t);
  out = normalized(-wi * eta + n * (eta * cos_i - cos_t));
  return true;
}
__device__ __forceinline__ float fresnel_dielectric(float eta_i, float eta_t, float cos_i) {  // ggx.rs:19-48
  cos_i = clampf(cos_i, -1.0f, 1.0f);
  if (cos_i < 0.0f) {
    cos_i = -cos_i;
    float tmp = eta_i;
    eta_i = eta_t;
    eta_t = tmp;
  }
  float sin_t = eta_i / eta_t * sqrtf(fmaxf(0.0f, 1.0f - cos_i * cos_i));
  float cos_t = sqrtf(fmaxf(0.0f, 1.0f - sin_t * sin_t));
  float ei_ct = eta_i * cos_t, et_ci = eta_t * cos_i, ei_ci = eta_i * cos_i, et_ct = eta_t * cos_t;
  float r_par = (et_ci - ei_ct) / (et_ci + ei_ct);
  float r_perp = (ei_ci - et_ct) / (ei_ci + et_ct);
  return (r_par * r_par + r_perp * r_perp) / 2.0f;
}
__device__ __forceinline__ float fresnel_conductor(float eta_i, float eta_t, float k_t, float cos_theta_i) {  // ggx.rs:50-85
  cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
  if (cos_theta_i < 0.0f) {
    cos_theta_i = -cos_theta_i;
    float tmp = eta_i;
    eta_i = eta_t;
    eta_t = tmp;
  }
  float eta = eta_t / eta_i, etak = k_t / eta_i;
  float c2 = cos_theta_i * cos_theta_i, s2 = 1.0f - c2;
  float eta2 = eta * eta, etak2 = etak * etak;
  float t0 = eta2 - etak2 - s2;
  float a2plusb2 = sqrtf(t0 * t0 + eta2 * etak2 * 4.0f);
  float t1 = a2plusb2 + c2;
  float a = sqrtf((a2plusb2 + t0) * 0.5f);
  float t2 = a * cos_theta_i * 2.0f;
  float rs = (t1 - t2) / (t1 + t2);
  float t3 = a2plusb2 * c2 + s2 * s2;
  float t4 = t2 * s2;
  float rp = rs * (t3 - t4) / (t3 + t4);
  return (rs + rp) / 2.0f;
}
__device__ __forceinline__ float ggx_d(float alpha, float3 wm) {  // ggx.rs:87-97
  float s0 = wm.x / alpha, s1 = wm.y / alpha;
  float t = wm.z * wm.z + s0 * s0 + s1 * s1;
  float a2 = alpha * alpha, t2 = t * t;
  return 1.0f / (RPT_PI * (a2 * t2));
}
__device__ __forceinline__ float ggx_lambda(float alpha, float3 w) {  // ggx.rs:99-107
  if (w.z == 0.0f) return 0.0f;
  float a2 = alpha * alpha;
  float c = 1.0f + (a2 * (w.x * w.x) + a2 * (w.y * w.y)) / (w.z * w.z);
  return sqrtf(c) * 0.5f - 0.5f;
}
__device__ __forceinline__ float ggx_g(float alpha, float3 wi, float3 wo) { return 1.0f / (1.0f + ggx_lambda(alpha, wi) + ggx_lambda(alpha, wo)); }
__device__ __forceinline__ float ggx_vnpdf(float alpha, float3 wi, float3 wh) {  // ggx.rs:115-119
  float inv_gl = 1.0f + ggx_lambda(alpha, wi);
  return (ggx_d(alpha, wh) * fabsf(dot(wi, wh))) / (inv_gl * fabsf(wi.z));
}
__device__ __forceinline__ float ggx_vnpdf_no_d(float alpha, float3 wi, float3 wh) {  // ggx.rs:121-123
  return fabsf(dot(wi, wh) / ((1.0f + ggx_lambda(alpha, wi)) * wi.z));
}
__device__ __forceinline__ float3 ggx_sample_vndf(float alpha, float3 wi, float x, float y) {  // ggx.rs:129-169
  float3 v = normalized(f3(alpha * wi.x, alpha * wi.y, wi.z));
  float3 t1 = v.z < 0.9999f ? normalized(cross(v, f3(0, 0, 1))) : f3(1, 0, 0);
  float3 t2 = cross(t1, v);
  float a = 1.0f / (1.0f + v.z);
  float r = sqrtf(x);
  float phi = y < a ? y / a * RPT_PI : RPT_PI + (y - a) / (1.0f - a) * RPT_PI;
  float sin_phi, cos_phi;
  sincosf(phi, &sin_phi, &cos_phi);
  float p1 = r * cos_phi;
  float p2 = r * sin_phi * (y < a ? 1.0f : v.z);
  float value = 1.0f - p1 * p1 - p2 * p2;
  float3 n = p1 * t1 + p2 * t2 + sqrtf(fmaxf(value, 0.0f)) * v;
  return normalized(f3(alpha * n.x, alpha * n.y, fmaxf(n.z, 0.0f)));
}
__device__ __forceinline__ float3 ggx_sample_wh(float alpha, float3 wi, float x, float y) {  // ggx.rs:171-180
  bool flip = wi.z < 0.0f;
  float3 wh = ggx_sample_vndf(alpha, flip ? -wi : wi, x, y);
  return flip ? -wh : wh;
}
struct GgxParams {
  float alpha, eta_inner, eta_outer, kappa;
  bool metallic;
};
__device__ __forceinline__ float ggx_reflectance(const GgxParams &g, float c) {  // ggx.rs:221-227
  return g.metallic ? fresnel_conductor(g.eta_outer, g.eta_inner, g.kappa, c) : fresnel_dielectric(g.eta_outer, g.eta_inner, c);
}
__device__ __forceinline__ float ggx_reflectance_probability(const GgxParams &g, float c) {  // ggx.rs:229-242
  return g.metallic ? 1.0f : clampf(ggx_reflectance(g, c), 0.0f, 1.0f);
}
__device__ __forceinline__ float ggx_eta_rel(const GgxParams &g, float3 wi) {  // ggx.rs:243-252
  return wi.z < 0.0f ? g.eta_outer / g.eta_inner : g.eta_inner / g.eta_outer;
}
// reflection lobe terms (ggx.rs:286-303 / 462-473)
__device__ __forceinline__ void ggx_reflect_terms(const GgxParams &g, float3 wi, float3 wo, float3 wh, float gcos, float ndotv, float &glossy,
                                                  float &glossy_pdf) {
  float refl = ggx_reflectance(g, ndotv);
  float d = ggx_d(g.alpha, wh);
  float gg = ggx_g(g.alpha, wi, wo);
  glossy = refl * (0.25f / gcos) * d * gg;
  glossy_pdf = fabsf(ndotv) == 0.0f ? 0.0f : ggx_vnpdf(g.alpha, wi, wh) * 0.25f / fabsf(ndotv);
}
// transmission lobe terms, TransportMode::Importance (ggx.rs:318-368 / 488-537)
__device__ __forceinline__ void ggx_transmit_terms(const GgxParams &g, float3 wi, float3 wo, float3 wh, float gcos, float &transmission,
                                                   float &transmission_pdf) {
  float eta_rel = ggx_eta_rel(g, wi);
  float gg = ggx_g(g.alpha, wi, wo);
  float partial = ggx_vnpdf_no_d(g.alpha, wi, wh);
  float ndotv = dot(wi, wh), ndotl = dot(wo, wh);
  float sqrt_denom = ndotv + eta_rel * ndotl;
  float eta_rel2 = eta_rel * eta_rel;
  float dwh_dwo1 = ndotl / (sqrt_denom * sqrt_denom);
  float dwh_dwo2 = eta_rel2 * dwh_dwo1;
  dwh_dwo1 = dwh_dwo2;  // Importance mode (ggx.rs:517-519)
  float d = ggx_d(g.alpha, wh);
  float weight = d * gg * ndotv * dwh_dwo1 / gcos;
  transmission_pdf = fabsf(d * partial * dwh_dwo2);
  float inv_reflectance = 1.0f - ggx_reflectance(g, ndotv);
  transmission = g.metallic ? 0.0f : inv_reflectance * fabsf(weight);
}
// GGX::bsdf (ggx.rs:256-400)
__device__ __forceinline__ Bsdf ggx_bsdf(const GgxParams &g, float3 wi, float3 wo) {
  wi = normalized(wi);
  bool same_hemisphere = wi.z * wo.z > 0.0f;
  float gcos = fabsf(wi.z * wo.z);
  Bsdf r;
  r.f = 0.0f;
  r.pdf = 0.0f;
  if (gcos == 0.0f) return r;
  float glossy = 0.0f, transmission = 0.0f, glossy_pdf = 0.0f, transmission_pdf = 0.0f;
  if (same_hemisphere) {
    float3 wh = normalized(wo + wi);
    if (wh.z < 0.0f) wh = -wh;
    ggx_reflect_terms(g, wi, wo, wh, gcos, dot(wi, wh), glossy, glossy_pdf);
  } else if (!g.metallic) {
    float eta_rel = ggx_eta_rel(g, wi);
    float3 wh = normalized(wi + eta_rel * wo);
    if (wh.z < 0.0f) wh = -wh;
    ggx_transmit_terms(g, wi, wo, wh, gcos, transmission, transmission_pdf);
  }
  float refl_prob = ggx_reflectance_probability(g, wi.z);  // evaluated at wi.z (reference quirk Q8)
  r.f = glossy + transmission;
  r.pdf = refl_prob * glossy_pdf + (1.0f - refl_prob) * transmission_pdf;
  return r;
}
// GGX::generate_and_evaluate (ggx.rs:401-590)
__device__ __forceinline__ Bsdf ggx_generate_and_evaluate(const GgxParams &g, float sx, float sy, float3 wi, float3 &wo) {
  float3 wh = normalized(ggx_sample_wh(g.alpha, wi, sx, sy));
  float refl_prob = ggx_reflectance_probability(g, dot(wh, wi));
  bool did_reflect = false;
  if (sx <= refl_prob) {  // the same sample.x that drove the VNDF radius (Q7)
    did_reflect = true;
    wo = ggx_reflect(wi, wh);
  } else {
    float eta_rel = 1.0f / ggx_eta_rel(g, wi);
    if (!ggx_refract(wi, wh, eta_rel, wo)) {
      did_reflect = true;
      wo = ggx_reflect(wi, wh);
    }
  }
  Bsdf r;
  r.f = 0.0f;
  r.pdf = 0.0f;
  float gcos = fabsf(wi.z * wo.z);
  if (gcos == 0.0f) return r;
  float cos_i;
  float glossy = 0.0f, transmission = 0.0f, glossy_pdf = 0.0f, transmission_pdf = 0.0f;
  if (did_reflect) {
    cos_i = dot(wi, wh);
    ggx_reflect_terms(g, wi, wo, wh, gcos, cos_i, glossy, glossy_pdf);
  } else {
    if (wh.z < 0.0f) wh = -wh;
    cos_i = dot(wi, wh);
    ggx_transmit_terms(g, wi, wo, wh, gcos, transmission, transmission_pdf);
  }
  float rp = ggx_reflectance_probability(g, cos_i);
  r.f = glossy + transmission;
  r.pdf = rp * glossy_pdf + (1.0f - rp) * transmission_pdf;
  return r;
}
__device__ __forceinline__ GgxParams ggx_params(const DevScene &S, const RptMaterial &m, float lambda) {
  GgxParams g;
  g.alpha = m.alpha;
  g.eta_inner = curve_eval(S, m.curve_a, lambda);
  g.eta_outer = curve_eval(S, m.curve_b, lambda);
  g.metallic = m.metallic != 0;
  g.kappa = g.metallic ? curve_eval(S, m.curve_c, lambda) : 0.0f;
  return g;
}

// Diffuse-like albedo: Lambertian texture (lambertian.rs:26) or a light's bounce_color (diffuse_light.rs:39).
__device__ __forceinline__ float diffuse_albedo(const DevScene &S, const RptMaterial &m, float lambda, float u, float v) {
  if (m.type == RPT_MATERIAL_LAMBERTIAN) return fminf(texstack_eval(S, m.texstack, lambda, u, v), 1.0f);
  return clampf(curve_eval(S, m.curve_a, lambda), 0.0f, 1.0f);
}
__device__ __forceinline__ float diffuse_albedo_fast(const DevScene &S, const RptMaterial &m, float2 fast, float lambda, float u, float v) {
  const int32_t c = __float_as_int(fast.x);
  if (m.type == RPT_MATERIAL_LAMBERTIAN && c >= 0) return fminf(0.0f + curve_eval(S, c, lambda) * fast.y, 1.0f);
  return diffuse_albedo(S, m, lambda, u, v);
}
// MaterialEnum::emission (diffuse_light.rs:123-133, sharp_light.rs:138-150,202-204); 0 for non-lights.
__device__ __forceinline__ float material_emission(const DevScene &S, const RptMaterial &m, float lambda, float3 wi) {
  if (m.type != RPT_MATERIAL_DIFFUSE_LIGHT && m.type != RPT_MATERIAL_SHARP_LIGHT) return 0.0f;
  float cosine = wi.z;
  bool ok = (cosine > 0.0f && m.sidedness == RPT_SIDED_FORWARD) || (cosine < 0.0f && m.sidedness == RPT_SIDED_REVERSE) ||
            m.sidedness == RPT_SIDED_DUAL;
  if (!ok) return 0.0f;
  float e = curve_eval(S, m.curve_b, lambda);
  if (m.type == RPT_MATERIAL_DIFFUSE_LIGHT) return e / RPT_PI;
  return e * ((m.sharpness + 1.0f) * powf(fabsf(wi.z), m.sharpness) / 2.0f / RPT_PI);
}

// ---------------------------------------------------------------------------------------------
// CurveWithCDF (Linear variant) — importance-map tables
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float nearest_curve_eval(const float *__restrict__ signal, uint32_t n, float x) {
  // Curve::Linear{bounds (0,1), Nearest}::evaluate
  if (x < 0.0f || x > 1.0f) return 0.0f;
  float step = 1.0f / (float)n;
  uint32_t index = (uint32_t)(x / step);
  if (index >= n) index = n - 1;
  float left = __ldg(signal + index);
  if (index + 1 >= n) return left;
  float t = (x - (float)index * step) / step;
  return t < 0.5f ? left : __ldg(signal + index + 1);
}
// Guide tables for the inversion below. guide[k] = first index i with cdf[i] >= k * (top / RPT_IMAP_GUIDE), top = the value the
// inversion scales its sample by. Builder (k_imap_guides) and sampler evaluate the thresholds with the same f32 expression, so
// the bracket [guide[k], guide[k + 1]] provably contains the index the full binary search returns (the CDF is non-decreasing).
#define RPT_IMAP_GUIDE 256u
__device__ __forceinline__ float imap_guide_threshold(uint32_t k, float top) { return (float)k * (top / (float)RPT_IMAP_GUIDE); }
__device__ __forceinline__ float nearest_curve_eval(const float *__restrict__ signal, uint32_t n, float x);
__device__ __forceinline__ uint32_t cdf_lower_bound(const float *__restrict__ cdf, uint32_t lo, uint32_t hi, float s) {
  while (lo < hi) {  // first index in [lo, hi) with cdf[i] >= s, else hi
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) < s) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ void nearest_cdf_sample(const float *__restrict__ pdf, const float *__restrict__ cdf, uint32_t n, float pdf_integral,
                                                   float sample, float &x_out, float &pdf_out, const uint32_t *__restrict__ guide = nullptr) {
  float lower_cdf = 0.0f;  // cdf.evaluate(-0.0001) is out of bounds
  float upper_cdf = nearest_curve_eval(cdf, n, 1.0f - 0.0001f);
  float s = lower_cdf + sample * (upper_cdf - lower_cdf);
  uint32_t lo = 0, hi = n;  // first index with cdf[i] >= s
  if (guide) {
    uint32_t k = min((uint32_t)(sample * (float)RPT_IMAP_GUIDE), RPT_IMAP_GUIDE - 1u);
    while (k > 0u && s < imap_guide_threshold(k, upper_cdf)) --k;                           // threshold(k) <= s
    while (k + 1u < RPT_IMAP_GUIDE && s > imap_guide_threshold(k + 1u, upper_cdf)) ++k;     // s <= threshold(k + 1), or the last bracket
    lo = __ldg(guide + k);
    if (k + 1u < RPT_IMAP_GUIDE) hi = __ldg(guide + k + 1u);
  }
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) < s) lo = mid + 1; else hi = mid;
  }
  uint32_t index = lo;
  float x;
  if (index == 0) {
    x = 0.0f;
  } else {
    if (index >= n) index = n - 1;
    float left = ((float)index - 1.0f) * 1.0f / (float)n;
    float right = (float)index * 1.0f / (float)n;
    float v0 = __ldg(cdf + index - 1), v1 = __ldg(cdf + index);
    float t = (s - v0) / (v1 - v0);
    x = t < 0.5f ? left : right;
  }
  x_out = x;
  pdf_out = nearest_curve_eval(pdf, n, x) / pdf_integral;
}

// ---------------------------------------------------------------------------------------------
// environment (world/environment.rs)
// ---------------------------------------------------------------------------------------------
// uv -> direction -> rotate -> uv (the HDR environment's lookups all go through it)
__device__ __forceinline__ float2 env_roundtrip(const DevScene &S, const float4 *rot, float u, float v) {
  if (S.env_unrotated) return uv_roundtrip_unrotated_cr(u, v);
  return direction_to_uv_cr(xform_vec(rot, uv_to_direction_cr(u, v)));
}
__device__ __forceinline__ bool env_in_sun(const DevScene &S, float u, float v) {
  float3 dir = uv_to_direction(u, v);
  float c = dot(S.env_sun_dir, dir);
  float s = sqrtf(1.0f - c * c);
  return fabsf(s) < sinf(S.env_angular_diameter / 2.0f) && c > 0.0f;
}
__device__ __forceinline__ float env_emission(const DevScene &S, float u, float v, float lambda) {  // :56-98
  if (S.env_kind == RPT_ENV_CONSTANT) return curve_eval(S, S.env_curve, lambda) * S.env_strength;
  if (S.env_kind == RPT_ENV_SUN) return env_in_sun(S, u, v) ? curve_eval(S, S.env_curve, lambda) * S.env_strength : 0.0f;
  float2 q = env_roundtrip(S, S.env_rot_rev, u, v);
  return texstack_eval(S, S.env_texstack, lambda, q.x, q.y) * S.env_strength;
}
__device__ __forceinline__ float env_pdf_for(const DevScene &S, float u, float v) {  // :198-258
  if (S.env_kind == RPT_ENV_CONSTANT) return 1.0f / (4.0f * RPT_PI);
  if (S.env_kind == RPT_ENV_SUN) return env_in_sun(S, u, v) ? 1.0f / (2.0f * RPT_PI * (1.0f - cosf(S.env_angular_diameter))) : 0.0f;
  if (S.imap_rows == 0) return 1.0f / (4.0f * RPT_PI);
  float2 q = env_roundtrip(S, S.env_rot_rev, u, v);
  float uu = q.x, vv = q.y;
  float m = nearest_curve_eval(S.imap_m_pdf, S.imap_marginal_n, uu);
  uint32_t row = (uint32_t)(clampf(uu, 0.0f, 1.0f - RPT_EPS) * (float)S.imap_rows);
  float r = nearest_curve_eval(S.imap_row_pdf + (size_t)row * S.imap_cols, S.imap_cols, vv);
  return m * r * (2.0f * RPT_PI * RPT_PI * sinf(RPT_PI * vv) + 0.001f) + 0.001f;
}
__device__ __forceinline__ void env_sample_uv(const DevScene &S, float sx, float sy, float &u, float &v, float &pdf) {  // :303-353
  if (S.env_kind == RPT_ENV_SUN) {
    float3 local_wo = f3(0, 0, 1) + sinf(S.env_angular_diameter / 2.0f) * random_in_unit_disk(sx, sy);
    Frame f = frame_from_normal(S.env_sun_dir);
    direction_to_uv(normalized(to_world(f, local_wo)), u, v);
    pdf = 1.0f / (2.0f * RPT_PI * (1.0f - cosf(S.env_angular_diameter)));
    return;
  }
  if (S.env_kind == RPT_ENV_CONSTANT || S.imap_rows == 0) {
    u = sx;
    v = sy;
    pdf = 1.0f / (4.0f * RPT_PI);
    return;
  }
  // ImportanceMap::sample_uv (importance_map.rs:325-357): sample.y -> row (u), sample.x -> column (v)
  float uu, row_pdf, vv, col_pdf;
  nearest_cdf_sample(S.imap_m_pdf, S.imap_m_cdf, S.imap_marginal_n, S.imap_marginal_integral, sy, uu, row_pdf, S.imap_m_guide);
  uint32_t row = (uint32_t)(uu * (float)S.imap_rows);
  if (row >= S.imap_rows) row = S.imap_rows - 1;
  nearest_cdf_sample(S.imap_row_pdf + (size_t)row * S.imap_cols, S.imap_row_cdf + (size_t)row * S.imap_cols, S.imap_cols, 1.0f, sx, vv, col_pdf,
                     S.imap_row_guide ? S.imap_row_guide + (size_t)row * RPT_IMAP_GUIDE : nullptr);
  float2 q = env_roundtrip(S, S.env_rot_fwd, uu, vv);
  u = q.x;
  v = q.y;
  pdf = row_pdf * col_pdf * (2.0f * RPT_PI * RPT_PI * sinf(RPT_PI * v) + 0.001f) + 0.001f;
}
