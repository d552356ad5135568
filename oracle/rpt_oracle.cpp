// rpt_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY: nothing in the product path
// (rust-pathtracer_b200/) may import, link or execute this file. It may be used by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, as the checker.
//
// A restatement, in plain scalar C++ (fp32, no FMA contraction: built with -ffp-contract=off),
// of the reference's PT path: gillett-hernandez/rust-pathtracer src/integrator/pt.rs and every
// per-ray function it calls. Each function cites the reference file:line it follows.
//
// PARITY PINNING. The reference cannot be built here (Rust nightly + two un-vendored git crates,
// SURVEY.md §8c) and ships no numeric golden vectors for this path. What pins this oracle:
//   * GGX sign/positivity properties of the reference's proptests + its saved regression case
//     (src/materials/ggx.rs:637-883, proptest-regressions/materials/ggx.txt:7)  -> tests/
//   * the 2-D importance-sampling integral 3.11227031972 within 1e-3
//     (src/world/importance_map.rs:798-942)                                      -> tests/
//   * tile coverage (src/renderer/tiled.rs:676-689), parse fixtures (data/test/*)
// Everything that lives in the un-vendored `math` crate (TangentFrame, random_cosine_direction,
// power_heuristic, PDF conversions, uv<->direction, CIE fits) is restated
// from the crate's published behaviour: **parity unpinned** for those items.
//
// Build: oracle/Makefile -> oracle/_build/librpt_oracle.so. Exported symbols are rpto_* with the
// same signatures as include/rpt.h's rpt_*.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/rpt.h"
#include "../include/rpt_rng.h"

namespace {

constexpr float PI_F = 3.14159265358979323846f;
constexpr float TAU_F = 6.28318530717958647692f;
constexpr float INF_F = std::numeric_limits<float>::infinity();
constexpr float EPS_F = 1.1920929e-7f;          // f32::EPSILON
constexpr float NORMAL_OFFSET = 0.001f;          // src/lib.rs:48
constexpr uint32_t NONE_U32 = 0xFFFFFFFFu;

thread_local std::string g_error;
bool g_debug = false;  // rpto_debug_color() turns on a per-sample trace (stdout)

// ---- math::Vec3 / Point3 (f32x4 newtypes; w handled implicitly) ---------------------------------
struct V3 {
  float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float norm_squared(V3 a) { return dot(a, a); }
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { return a / norm(a); }
inline float signum(float x) { return std::isnan(x) ? x : (std::signbit(x) ? -1.0f : 1.0f); }  // f32::signum
inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

struct Ray {
  V3 o, d;
  float tmax;
};

struct Mat4 {
  float m[16];
};
inline V3 mul_point(const Mat4 &M, V3 p) {
  return {M.m[0] * p.x + M.m[1] * p.y + M.m[2] * p.z + M.m[3], M.m[4] * p.x + M.m[5] * p.y + M.m[6] * p.z + M.m[7],
          M.m[8] * p.x + M.m[9] * p.y + M.m[10] * p.z + M.m[11]};
}
inline V3 mul_vec(const Mat4 &M, V3 v) {
  return {M.m[0] * v.x + M.m[1] * v.y + M.m[2] * v.z, M.m[4] * v.x + M.m[5] * v.y + M.m[6] * v.z,
          M.m[8] * v.x + M.m[9] * v.y + M.m[10] * v.z};
}
inline V3 mul_vec_transposed(const Mat4 &M, V3 v) {  // (M^T) * v, upper 3x3
  return {M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z, M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z,
          M.m[2] * v.x + M.m[6] * v.y + M.m[10] * v.z};
}

// ---- math::TangentFrame::from_normal (Duff et al. 2017; SURVEY Appendix B, unpinned) ------------
struct Frame {
  V3 t, b, n;
};
inline Frame frame_from_normal(V3 n) {
  float sign = std::copysign(1.0f, n.z);
  float a = -1.0f / (sign + n.z);
  float b = n.x * n.y * a;
  Frame f;
  f.t = v3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
  f.b = v3(b, sign + n.y * n.y * a, -n.y);
  f.n = n;
  return f;
}
inline V3 to_local(const Frame &f, V3 v) { return v3(dot(f.t, v), dot(f.b, v), dot(f.n, v)); }
inline V3 to_world(const Frame &f, V3 v) { return f.t * v.x + f.b * v.y + f.n * v.z; }

// ---- math sampling helpers (unpinned) ----------------------------------------
inline V3 random_cosine_direction(float sx, float sy) {
  float phi = TAU_F * sx;
  float r = std::sqrt(sy);
  float z = std::sqrt(1.0f - sy);
  return v3(std::cos(phi) * r, std::sin(phi) * r, z);
}
inline V3 random_on_unit_sphere(float sx, float sy) {
  float phi = sx * TAU_F;
  float z = sy * 2.0f - 1.0f;
  float r = std::sqrt(1.0f - z * z);
  return v3(r * std::cos(phi), r * std::sin(phi), z);
}
inline V3 random_in_unit_disk(float sx, float sy) {
  float u = sx * TAU_F;
  float v = std::sqrt(sy);
  return v3(std::cos(u) * v, std::sin(u) * v, 0.0f);
}
inline V3 uv_to_direction(float u, float v) {
  float theta = (u - 0.5f) * TAU_F;
  float phi = v * PI_F;
  float st = std::sin(theta), ctt = std::cos(theta), sp = std::sin(phi), cp = std::cos(phi);
  return v3(sp * ctt, sp * st, cp);
}
inline void direction_to_uv(V3 d, float &u, float &v) {
  float theta = std::atan2(d.y, d.x);
  float phi = std::acos(d.z);
  u = theta / 2.0f / PI_F + 0.5f;
  v = phi / PI_F;
}
// The same maps with correctly rounded f32 trigonometry (through f64), used by the HDR environment only. The reference
// calls the platform libm (f32::sin_cos / atan2 / acos); importance-map samples sit exactly on texel boundaries, so which
// texel the round trip uv -> direction -> uv lands in depends on the libm's last ulp. The oracle and the CUDA path both
// pin the correctly rounded value (what glibc >= 2.41 returns), see DESIGN.md §3.
inline V3 uv_to_direction_cr(float u, float v) {
  float theta = (u - 0.5f) * TAU_F;
  float phi = v * PI_F;
  float st = (float)std::sin((double)theta), ctt = (float)std::cos((double)theta), sp = (float)std::sin((double)phi), cp = (float)std::cos((double)phi);
  return v3(sp * ctt, sp * st, cp);
}
inline void direction_to_uv_cr(V3 d, float &u, float &v) {
  float theta = (float)std::atan2((double)d.y, (double)d.x);
  float phi = (float)std::acos((double)d.z);
  u = theta / 2.0f / PI_F + 0.5f;
  v = phi / PI_F;
}
inline float power_heuristic(float a, float b) { return (a * a) / (a * a + b * b); }
inline float power_heuristic_generic(float a, float b) { return a / (a + b); }  // src/lib.rs:114-119
// Sample1D::choose(split, a, b)
inline bool choose(float x, float split, float &rescaled) {  // returns true for the first option
  if (x < split) {
    rescaled = clampf(x / split, 0.0f, 1.0f - EPS_F);
    return true;
  }
  rescaled = clampf((x - split) / (1.0f - split), 0.0f, 1.0f - EPS_F);
  return false;
}

// ---- AABB (src/aabb.rs) ------------------------------------------------------------------------
struct AABB {
  V3 mn, mx;
};
inline AABB aabb_empty() { return {v3(INF_F, INF_F, INF_F), v3(-INF_F, -INF_F, -INF_F)}; }
inline V3 vmin(V3 a, V3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline AABB aabb_new(V3 a, V3 b) { return {vmin(a, b), vmax(a, b)}; }           // aabb.rs:16-21
inline AABB aabb_expand(AABB a, const AABB &o) { return {vmin(a.mn, o.mn), vmax(a.mx, o.mx)}; }
inline AABB aabb_grow(AABB a, V3 p) { return {vmin(a.mn, p), vmax(a.mx, p)}; }
inline V3 aabb_size(const AABB &a) { return a.mx - a.mn; }
inline V3 aabb_center(const AABB &a) { return a.mn + aabb_size(a) / 2.0f; }  // aabb.rs:95-97
inline float aabb_surface_area(const AABB &a) {                              // aabb.rs:99-102
  V3 s = aabb_size(a);
  return 2.0f * (s.x * s.y + s.x * s.z + s.y * s.z);
}
// aabb.rs:37-65 called with (t0, t1) = (0, +inf) — the only way the reference calls it (F8).
inline bool aabb_hit(const AABB &b, const Ray &r) {
  const float o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
  const float mn[3] = {b.mn.x, b.mn.y, b.mn.z}, mx[3] = {b.mx.x, b.mx.y, b.mx.z};
  float tmin_max = 0.0f;   // w lane: direction.w == 0 -> (0, inf)
  float tmax_min = INF_F;
  bool any_tmax_neg = false;
  for (int k = 0; k < 3; ++k) {
    float lo = d[k] == 0.0f ? 0.0f : (mn[k] - o[k]) / d[k];
    float hi = d[k] == 0.0f ? INF_F : (mx[k] - o[k]) / d[k];
    float tmin = std::fmin(lo, hi), tmax = std::fmax(lo, hi);
    tmin_max = std::fmax(tmin_max, tmin);
    tmax_min = std::fmin(tmax_min, tmax);
    if (tmax < 0.0f) any_tmax_neg = true;  // tmax.simd_lt(scaled_t0 = 0).any()
  }
  if (tmin_max > tmax_min) return false;
  if (any_tmax_neg) return false;
  return true;
}
// Matrix4x4 * AABB (aabb.rs:116-138)
inline AABB aabb_transform(const Mat4 &M, const AABB &b) {
  V3 mn = v3(INF_F, INF_F, INF_F), mx = v3(-INF_F, -INF_F, -INF_F);
  for (int i = 0; i < 8; ++i) {
    bool xb = (i & 1) == 0, yb = ((i >> 1) & 1) == 0, zb = ((i >> 2) & 1) == 0;
    V3 c = mul_point(M, v3(xb ? b.mn.x : b.mx.x, yb ? b.mn.y : b.mx.y, zb ? b.mn.z : b.mx.z));
    mn = vmin(mn, c);
    mx = vmax(mx, c);
  }
  return {mn, mx};
}

// ---- BVH build (src/accelerator/bvh.rs:299-457) + flatten (lbvh.rs:47-134) -----------------------
struct BNode {
  bool leaf;
  uint32_t shape;
  uint32_t l, r;
  AABB la, ra;
};
struct FlatNode {  // lbvh.rs:16-45
  AABB aabb;
  uint32_t entry, exit, shape;
};
struct Bucket {
  size_t size;
  AABB aabb;
};

struct BvhBuilder {
  const std::vector<AABB> &shapes;
  std::vector<BNode> nodes;
  explicit BvhBuilder(const std::vector<AABB> &s) : shapes(s) {}

  AABB joint(const std::vector<uint32_t> &idx, size_t b, size_t e) {
    AABB a = aabb_empty();
    for (size_t i = b; i < e; ++i) a = aabb_expand(a, shapes[idx[i]]);
    return a;
  }

  uint32_t build(const std::vector<uint32_t> &indices) {
    AABB bounds = aabb_empty(), cbounds = aabb_empty();
    for (uint32_t i : indices) {
      bounds = aabb_expand(bounds, shapes[i]);
      cbounds = aabb_grow(cbounds, aabb_center(shapes[i]));
    }
    if (indices.size() == 1) {
      nodes.push_back(BNode{true, indices[0], 0, 0, aabb_empty(), aabb_empty()});
      return (uint32_t)nodes.size() - 1;
    }
    uint32_t me = (uint32_t)nodes.size();
    nodes.push_back(BNode{true, 0, 0, 0, aabb_empty(), aabb_empty()});  // dummy
    V3 size = aabb_size(cbounds);
    float sz[3] = {size.x, size.y, size.z};
    float max_axis = std::fmax(std::fmax(sz[0], sz[1]), std::fmax(sz[2], 0.0f));  // w lane size == 0
    int split_axis = 3;                                                          // w lane wins only if max == 0
    if (!(0.0f >= max_axis)) {
      split_axis = 0;
      for (int k = 0; k < 3; ++k)
        if (sz[k] >= max_axis) split_axis = k;  // mask.select([0,1,2,3], 0).reduce_max()
    }
    const float cmn[3] = {cbounds.mn.x, cbounds.mn.y, cbounds.mn.z};
    float split_axis_size = split_axis == 3 ? 0.0f : sz[split_axis];
    uint32_t li, ri;
    AABB la, ra;
    if (split_axis_size < 0.00001f) {
      size_t half = indices.size() / 2;
      std::vector<uint32_t> L(indices.begin(), indices.begin() + half), R(indices.begin() + half, indices.end());
      la = joint(L, 0, L.size());
      ra = joint(R, 0, R.size());
      li = build(L);
      ri = build(R);
    } else {
      constexpr int NB = 6;
      Bucket buckets[NB];
      std::vector<uint32_t> assign[NB];
      for (auto &b : buckets) b = Bucket{0, aabb_empty()};
      for (uint32_t idx : indices) {
        V3 c = aabb_center(shapes[idx]);
        const float cc[3] = {c.x, c.y, c.z};
        float rel = (cc[split_axis] - cmn[split_axis]) / split_axis_size;
        size_t bn = (size_t)(rel * ((float)NB - 0.01f));
        if (bn >= NB) bn = NB - 1;  // unreachable in the reference (would panic); guards fp edge
        buckets[bn].size += 1;
        buckets[bn].aabb = aabb_expand(buckets[bn].aabb, shapes[idx]);
        assign[bn].push_back(idx);
      }
      int min_bucket = 0;
      float min_cost = INF_F;
      la = aabb_empty();
      ra = aabb_empty();
      for (int i = 0; i < NB - 1; ++i) {
        Bucket l{0, aabb_empty()}, r{0, aabb_empty()};
        for (int k = 0; k <= i; ++k) l = Bucket{l.size + buckets[k].size, aabb_expand(l.aabb, buckets[k].aabb)};
        for (int k = i + 1; k < NB; ++k) r = Bucket{r.size + buckets[k].size, aabb_expand(r.aabb, buckets[k].aabb)};
        float cost = ((float)l.size * aabb_surface_area(l.aabb) + (float)r.size * aabb_surface_area(r.aabb)) /
                     aabb_surface_area(bounds);
        if (cost < min_cost) {  // NaN (empty side: 0 * inf) never wins
          min_bucket = i;
          min_cost = cost;
          la = l.aabb;
          ra = r.aabb;
        }
      }
      std::vector<uint32_t> L, R;
      for (int k = 0; k <= min_bucket; ++k) L.insert(L.end(), assign[k].begin(), assign[k].end());
      for (int k = min_bucket + 1; k < NB; ++k) R.insert(R.end(), assign[k].begin(), assign[k].end());
      li = build(L);
      ri = build(R);
    }
    nodes[me] = BNode{false, 0, li, ri, la, ra};
    return me;
  }

  void flatten_branch(uint32_t node, const AABB &aabb, std::vector<FlatNode> &out) {
    size_t me = out.size();
    out.push_back(FlatNode{aabb_empty(), 0, 0, 0});
    flatten(node, out);
    out[me] = FlatNode{aabb, (uint32_t)me + 1, (uint32_t)out.size(), NONE_U32};
  }
  void flatten(uint32_t node, std::vector<FlatNode> &out) {
    const BNode n = nodes[node];
    if (n.leaf) {
      uint32_t next = (uint32_t)out.size() + 1;
      out.push_back(FlatNode{aabb_empty(), NONE_U32, next, n.shape});
    } else {
      flatten_branch(n.l, n.la, out);
      flatten_branch(n.r, n.ra, out);
    }
  }
};

std::vector<FlatNode> build_flat_bvh(const std::vector<AABB> &shapes) {
  std::vector<FlatNode> out;
  if (shapes.empty()) return out;
  BvhBuilder b(shapes);
  std::vector<uint32_t> idx(shapes.size());
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
  b.build(idx);
  b.flatten(0, out);
  return out;
}

// FlatBVH::traverse (lbvh.rs:172-213): every leaf whose own AABB the forward ray touches, DFS order.
template <class F>
inline void bvh_traverse(const std::vector<FlatNode> &bvh, const std::vector<AABB> &shape_aabbs, const Ray &r, F &&visit) {
  size_t index = 0, n = bvh.size();
  while (index < n) {
    const FlatNode &node = bvh[index];
    if (node.entry == NONE_U32) {
      if (aabb_hit(shape_aabbs[node.shape], r)) visit(node.shape);
      index = node.exit;
    } else if (aabb_hit(node.aabb, r)) {
      index = node.entry;
    } else {
      index = node.exit;
    }
  }
}

// ---- scene ---------------------------------------------------------------------------------------
struct Hit {  // src/hittable.rs:7-16 (+ primitive id for the parity hook)
  float t;
  V3 p;
  float u, v;
  V3 n;
  uint32_t material;
  uint32_t instance;
  uint32_t prim;
};

struct OMesh {
  std::vector<V3> verts, normals;
  std::vector<uint32_t> idx, fmat;
  std::vector<AABB> tri_aabb;
  std::vector<FlatNode> bvh;
  AABB bbox;
};

struct OTexture {
  uint32_t channels, w, h;
  std::vector<float> texels;
  int32_t curves[4];
};

}  // namespace

struct RptScene {  // oracle flavour of the opaque handle
  std::vector<RptInstance> instances;
  std::vector<OMesh> meshes;
  std::vector<AABB> inst_aabb;
  std::vector<FlatNode> tlas;
  std::vector<uint32_t> lights;
  std::vector<RptMaterial> materials;
  uint32_t num_lambda;
  float lut_lo, lut_hi;
  std::vector<float> curve_lut, cie_lut;
  std::vector<OTexture> textures;
  std::vector<uint32_t> stack_tex;
  std::vector<RptTexStack> stacks;
  RptEnvironment env;
  std::vector<float> imap_row_pdf, imap_row_cdf, imap_m_pdf, imap_m_cdf;
  float p_env;
  std::vector<RptCamera> cameras;
  float world_radius;
};

namespace {

inline V3 shuffle_axis(V3 v, uint32_t axis) {  // rect.rs:6-12
  switch (axis) {
    case RPT_AXIS_X: return v3(v.z, v.y, v.x);
    case RPT_AXIS_Y: return v3(v.x, v.z, v.y);
    default: return v;
  }
}
inline V3 axis_vec(uint32_t axis) { return axis == RPT_AXIS_X ? v3(1, 0, 0) : (axis == RPT_AXIS_Y ? v3(0, 1, 0) : v3(0, 0, 1)); }
inline V3 inst_origin(const RptInstance &I) { return v3(I.origin[0], I.origin[1], I.origin[2]); }
inline const Mat4 &fwd(const RptInstance &I) { return *reinterpret_cast<const Mat4 *>(I.forward); }
inline const Mat4 &rev(const RptInstance &I) { return *reinterpret_cast<const Mat4 *>(I.reverse); }

// -- curve LUT evaluation: the boundary contract of include/rpt.h (linear interpolation) -------------
inline float lut_eval(const float *lut, uint32_t n, float lo, float hi, float lambda) {
  float x = (lambda - lo) / (hi - lo) * (float)(n - 1);
  x = clampf(x, 0.0f, (float)(n - 1));
  uint32_t i = (uint32_t)x;
  if (i > n - 2) i = n - 2;
  float t = x - (float)i;
  float a = lut[i], b = lut[i + 1];
  return a + t * (b - a);
}
inline float curve_eval(const RptScene &S, int32_t c, float lambda) {
  return lut_eval(&S.curve_lut[(size_t)c * S.num_lambda], S.num_lambda, S.lut_lo, S.lut_hi, lambda);
}

// Texture{1,4}::eval_at + TexStack::eval_at (texture.rs:101-116,134-142,258-266; vec2d.rs:34-42)
inline float texstack_eval(const RptScene &S, int32_t stack, float lambda, float u, float v) {
  float energy = 0.0f;
  const RptTexStack &st = S.stacks[stack];
  for (uint32_t k = 0; k < st.count; ++k) {
    const OTexture &T = S.textures[S.stack_tex[st.first + k]];
    float uu = clampf(u, 0.0f, 1.0f - EPS_F), vv = clampf(v, 0.0f, 1.0f - EPS_F);
    size_t x = (size_t)(uu * (float)T.w), y = (size_t)(vv * (float)T.h);
    const float *tx = &T.texels[(y * T.w + x) * T.channels];
    if (T.channels == 1) {
      energy += curve_eval(S, T.curves[0], lambda) * tx[0];
    } else {
      float s = 0.0f;
      for (int c = 0; c < 4; ++c) s += curve_eval(S, T.curves[c], lambda) * tx[c];
      energy += s;
    }
  }
  return energy;
}

// ---- primitives ------------------------------------------------------------------------------------
// AARect::hit (rect.rs:69-112)
inline bool rect_hit(const RptInstance &I, const Ray &r, float t0, float t1, Hit &h) {
  V3 tmp_o = shuffle_axis(r.o - inst_origin(I), I.axis);
  V3 tmp_d = shuffle_axis(r.d, I.axis);
  if (tmp_d.z == 0.0f) return false;
  float t = (-tmp_o.z) / tmp_d.z;
  if (t <= t0 || t > t1 || t >= r.tmax) return false;
  float xh = tmp_o.x + t * tmp_d.x, yh = tmp_o.y + t * tmp_d.y;
  float hx = I.size[0] / 2.0f, hy = I.size[1] / 2.0f;
  if (xh < -hx || xh > hx || yh < -hy || yh > hy) return false;
  V3 n = axis_vec(I.axis);
  if (I.two_sided && dot(r.d, n) > 0.0f) n = -n;
  h.t = t;
  h.p = r.o + r.d * t;
  h.u = (xh + hx) / I.size[0];
  h.v = (yh + hy) / I.size[1];
  h.n = normalized(n);
  h.material = RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
  h.prim = 0;
  return true;
}
// Sphere::hit (sphere.rs:34-87)
inline bool sphere_hit(const RptInstance &I, const Ray &r, float t0, float t1, Hit &h) {
  float radius = I.size[0];
  V3 oc = r.o - inst_origin(I);
  float a = dot(r.d, r.d), b = dot(oc, r.d), c = dot(oc, oc) - radius * radius;
  float disc = b * b - a * c;
  float ds = std::sqrt(disc);
  if (disc > 0.0f) {
    for (int k = 0; k < 2; ++k) {
      float time = k == 0 ? (-b - ds) / a : (-b + ds) / a;
      if (time < t1 && time > t0 && time < r.tmax) {
        V3 p = r.o + r.d * time;
        h.t = time;
        h.p = p;
        h.u = h.v = 0.0f;
        h.n = normalized((p - inst_origin(I)) / radius);
        h.material = RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
        h.prim = 0;
        return true;
      }
    }
  }
  return false;
}
// Disk::hit (disk.rs:31-62)
inline bool disk_hit(const RptInstance &I, const Ray &r, float t0, float t1, Hit &h) {
  float radius = I.size[0];
  V3 tmp_o = r.o - inst_origin(I);
  V3 tmp_d = r.d;
  if (tmp_d.z == 0.0f) return false;
  float t = (-tmp_o.z) / tmp_d.z;
  if (t <= t0 || t > t1 || t >= r.tmax) return false;
  float xh = tmp_o.x + t * tmp_d.x, yh = tmp_o.y + t * tmp_d.y;
  if (xh * xh + yh * yh > radius * radius) return false;
  V3 n = v3(0, 0, 1);
  if (dot(r.d, n) > 0.0f && I.two_sided) n = -n;
  h.t = t;
  h.p = r.o + r.d * t;
  h.u = h.v = 0.0f;
  h.n = n;
  h.material = RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
  h.prim = 0;
  return true;
}

inline V3 tri_shuffle(V3 v, uint32_t kz) {  // mesh.rs:12-19
  switch (kz) {
    case 0: return v3(v.y, v.z, v.x);
    case 1: return v3(v.z, v.x, v.y);
    default: return v;
  }
}
// MeshTriangleRef::hit (mesh.rs:67-198)
inline bool tri_hit(const OMesh &M, uint32_t tri, const Ray &r, float t0, float t1, Hit &h) {
  uint32_t i0 = M.idx[3 * tri], i1 = M.idx[3 * tri + 1], i2 = M.idx[3 * tri + 2];
  V3 p0 = M.verts[i0], p1 = M.verts[i1], p2 = M.verts[i2];
  V3 p0t = p0 - r.o, p1t = p1 - r.o, p2t = p2 - r.o;
  float ax = std::fabs(r.d.x), ay = std::fabs(r.d.y), az = std::fabs(r.d.z);
  float mx = std::fmax(std::fmax(ax, ay), std::fmax(az, 0.0f));
  uint32_t kz = 0;
  if (ax >= mx) kz = 0;
  if (ay >= mx) kz = 1;
  if (az >= mx) kz = 2;
  if (0.0f >= mx) kz = 3;
  V3 d = tri_shuffle(r.d, kz);
  p0t = tri_shuffle(p0t, kz);
  p1t = tri_shuffle(p1t, kz);
  p2t = tri_shuffle(p2t, kz);
  float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
  p0t.x += sx * p0t.z;
  p1t.x += sx * p1t.z;
  p2t.x += sx * p2t.z;
  p0t.y += sy * p0t.z;
  p1t.y += sy * p1t.z;
  p2t.y += sy * p2t.z;
  float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    e0 = (float)((double)p2t.y * (double)p1t.x - (double)p2t.x * (double)p1t.y);
    e1 = (float)((double)p0t.y * (double)p2t.x - (double)p0t.x * (double)p2t.y);
    e2 = (float)((double)p1t.y * (double)p0t.x - (double)p1t.x * (double)p0t.y);
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = e0 + e1 + e2;
  if (det == 0.0f) return false;
  p0t.z *= sz;
  p1t.z *= sz;
  p2t.z *= sz;
  float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if ((det < 0.0f && (t_scaled >= t0 * det || t_scaled < t1 * det)) ||
      (det > 0.0f && (t_scaled <= t0 * det || t_scaled > t1 * det)))
    return false;
  float inv_det = 1.0f / det;
  float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
  V3 gn = normalized(cross(p0 - p2, p1 - p2));
  V3 n = gn;
  if (!M.normals.empty()) n = b0 * M.normals[i0] + b1 * M.normals[i1] + b2 * M.normals[i2];
  h.t = t_scaled * inv_det;
  h.p = b0 * p0 + b1 * p1 + b2 * p2;
  h.u = h.v = 0.0f;
  h.n = normalized(n);  // HitRecord::new normalises (hittable.rs:34)
  h.material = M.fmat.empty() ? RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0) : M.fmat[tri];
  h.prim = tri;
  return true;
}
// Mesh::hit (mesh.rs:314-360)
inline bool mesh_hit(const OMesh &M, const Ray &r, float t0, float t1, Hit &h) {
  float closest = t1;
  bool any = false;
  bvh_traverse(M.bvh, M.tri_aabb, r, [&](uint32_t tri) {
    Hit tmp;
    if (tri_hit(M, tri, r, t0, closest, tmp)) {
      closest = tmp.t;
      h = tmp;
      any = true;
    }
  });
  return any;
}

inline bool aggregate_hit(const RptScene &S, const RptInstance &I, const Ray &r, float t0, float t1, Hit &h) {
  switch (I.kind) {
    case RPT_AGG_RECT: return rect_hit(I, r, t0, t1, h);
    case RPT_AGG_SPHERE: return sphere_hit(I, r, t0, t1, h);
    case RPT_AGG_DISK: return disk_hit(I, r, t0, t1, h);
    default: return mesh_hit(S.meshes[I.mesh], r, t0, t1, h);
  }
}
// Instance::hit (instance.rs:75-133)
inline bool instance_hit(const RptScene &S, uint32_t id, const Ray &r, float t0, float t1, Hit &h) {
  const RptInstance &I = S.instances[id];
  if (I.has_transform) {
    Ray lr{mul_point(rev(I), r.o), mul_vec(rev(I), r.d), r.tmax};
    if (!aggregate_hit(S, I, lr, t0, t1, h)) return false;
    h.n = normalized(mul_vec_transposed(rev(I), h.n));
    h.p = mul_point(fwd(I), h.p);
  } else if (!aggregate_hit(S, I, r, t0, t1, h)) {
    return false;
  }
  h.instance = id;
  if (I.material != RPT_MAT_NONE) h.material = I.material;
  return true;
}
// World::hit -> Accelerator::hit, BVH arm (world/mod.rs:166, accelerator/mod.rs:107-176)
inline bool world_hit(const RptScene &S, const Ray &r, float t0, float t1, Hit &h) {
  float closest = t1;
  bool any = false;
  bvh_traverse(S.tlas, S.inst_aabb, r, [&](uint32_t id) {
    Hit tmp;
    if (instance_hit(S, id, r, t0, closest, tmp)) {
      closest = tmp.t;
      h = tmp;
      any = true;
    }
  });
  return any;
}

// ---- light sampling (Hittable::sample / psa_pdf) -----------------------------------------------------
inline float area_to_solid_angle(float p, float cos_i, float d2) { return p * d2 / std::fabs(cos_i); }
// AARect::sample_surface + sample (rect.rs:113-155)
inline void rect_sample(const RptInstance &I, float sx, float sy, V3 from, V3 &dir, float &pdf) {
  V3 n = axis_vec(I.axis);
  float x = sx;
  if (I.two_sided) {
    float resc;
    bool first = choose(x, 0.5f, resc);
    x = resc;
    n = n * (first ? -1.0f : 1.0f);
  }
  V3 point = inst_origin(I) + shuffle_axis(v3((x - 0.5f) * I.size[0], (sy - 0.5f) * I.size[1], 0.0f), I.axis);
  float area = I.size[0] * I.size[1];
  V3 direction = point - from;
  float cos_i = dot(n, normalized(direction));
  float p = area_to_solid_angle(1.0f / area, cos_i, norm_squared(direction));
  dir = normalized(direction);
  pdf = std::isfinite(p) ? p : 0.0f;
}
// Sphere::sample (sphere.rs:88-131)
inline void sphere_sample(const RptInstance &I, float sx, float sy, V3 from, V3 &dir, float &pdf) {
  float radius = I.size[0];
  V3 n = random_on_unit_sphere(sx, sy);
  V3 point = inst_origin(I) + radius * n;
  float area_pdf = 1.0f / (radius * radius * 4.0f * PI_F);
  V3 direction = point - from;
  float ndd = std::fabs(dot(n, normalized(direction)));
  float p = area_pdf * norm_squared(direction) / ndd;
  dir = normalized(direction);
  pdf = std::isfinite(p) ? p : 0.0f;
}
// Disk::sample (disk.rs:63-91)
inline void disk_sample(const RptInstance &I, float sx, float sy, V3 from, V3 &dir, float &pdf) {
  float radius = I.size[0];
  V3 n = v3(0, 0, 1);
  float x = sx;
  if (I.two_sided) {
    float resc;
    bool first = choose(x, 0.5f, resc);
    x = resc;
    n = n * (first ? -1.0f : 1.0f);
  }
  V3 point = inst_origin(I) + radius * random_in_unit_disk(x, sy);
  float area = PI_F * radius * radius;
  V3 direction = point - from;
  float cos_i = dot(n, normalized(direction));
  float p = area_to_solid_angle(1.0f / area, cos_i, norm_squared(direction));
  dir = normalized(direction);
  pdf = std::isfinite(p) ? p : 0.0f;
}
// Instance::sample (instance.rs:134-141)
inline bool instance_sample(const RptScene &S, uint32_t id, float sx, float sy, V3 from, V3 &dir, float &pdf) {
  const RptInstance &I = S.instances[id];
  V3 f = I.has_transform ? mul_point(rev(I), from) : from;
  switch (I.kind) {
    case RPT_AGG_RECT: rect_sample(I, sx, sy, f, dir, pdf); break;
    case RPT_AGG_SPHERE: sphere_sample(I, sx, sy, f, dir, pdf); break;
    case RPT_AGG_DISK: disk_sample(I, sx, sy, f, dir, pdf); break;
    default: return false;  // mesh lights: todo!() in the reference (mesh.rs:362-386)
  }
  if (I.has_transform) dir = normalized(mul_vec(fwd(I), dir));
  return true;
}
// Instance::psa_pdf + per-primitive psa_pdf (instance.rs:154-170, rect.rs:156-173, sphere.rs:132-152, disk.rs:92-104)
inline float instance_psa_pdf(const RptScene &S, uint32_t id, float cos_o, float cos_i, V3 from, V3 to) {
  const RptInstance &I = S.instances[id];
  if (I.has_transform) {  // to_world, not to_local (Q11)
    from = mul_point(fwd(I), from);
    to = mul_point(fwd(I), to);
  }
  float d2 = norm_squared(to - from);
  switch (I.kind) {
    case RPT_AGG_RECT: {
      float area_pdf = 1.0f / (I.size[0] * I.size[1]);
      return area_to_solid_angle(area_pdf, cos_i, d2) / std::fabs(cos_o);
    }
    case RPT_AGG_SPHERE: {
      float area_pdf = 1.0f / (I.size[0] * I.size[0] * 4.0f * PI_F);
      return area_pdf * d2 / std::fabs(cos_i * cos_o);
    }
    case RPT_AGG_DISK: {
      float area = PI_F * I.size[0] * I.size[0];
      return d2 / ((std::fabs(cos_o) * std::fabs(cos_i) + 0.00001f) * area);
    }
    default: return 0.0f;
  }
}

// ---- materials ----------------------------------------------------------------------------------------
struct Bsdf {
  float f, pdf;
};
inline V3 reflect(V3 wi, V3 n) {  // ggx.rs:3-6
  V3 w = -wi;
  return normalized(w - 2.0f * dot(w, n) * n);
}
inline bool refract(V3 wi, V3 n, float eta, V3 &out) {  // ggx.rs:8-17
  float cos_i = dot(wi, n);
  float sin2i = std::fmax(1.0f - cos_i * cos_i, 0.0f);
  float sin2t = eta * eta * sin2i;
  if (sin2t >= 1.0f) return false;
  float cos_t = std::sqrt(1.0f - sin2t);
  out = normalized(-wi * eta + n * (eta * cos_i - cos_t));
  return true;
}
inline float fresnel_dielectric(float eta_i, float eta_t, float cos_i) {  // ggx.rs:19-48
  cos_i = clampf(cos_i, -1.0f, 1.0f);
  if (cos_i < 0.0f) {
    cos_i = -cos_i;
    std::swap(eta_i, eta_t);
  }
  float sin_t = eta_i / eta_t * std::sqrt(std::fmax(0.0f, 1.0f - cos_i * cos_i));
  float cos_t = std::sqrt(std::fmax(0.0f, 1.0f - sin_t * sin_t));
  float ei_ct = eta_i * cos_t, et_ci = eta_t * cos_i, ei_ci = eta_i * cos_i, et_ct = eta_t * cos_t;
  float r_par = (et_ci - ei_ct) / (et_ci + ei_ct);
  float r_perp = (ei_ci - et_ct) / (ei_ci + et_ct);
  return (r_par * r_par + r_perp * r_perp) / 2.0f;
}
inline float fresnel_conductor(float eta_i, float eta_t, float k_t, float cos_theta_i) {  // ggx.rs:50-85
  cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
  if (cos_theta_i < 0.0f) {
    cos_theta_i = -cos_theta_i;
    std::swap(eta_i, eta_t);
  }
  float eta = eta_t / eta_i, etak = k_t / eta_i;
  float c2 = cos_theta_i * cos_theta_i, s2 = 1.0f - c2;
  float eta2 = eta * eta, etak2 = etak * etak;
  float t0 = eta2 - etak2 - s2;
  float a2plusb2 = std::sqrt(t0 * t0 + eta2 * etak2 * 4.0f);
  float t1 = a2plusb2 + c2;
  float a = std::sqrt((a2plusb2 + t0) * 0.5f);
  float t2 = a * cos_theta_i * 2.0f;
  float rs = (t1 - t2) / (t1 + t2);
  float t3 = a2plusb2 * c2 + s2 * s2;
  float t4 = t2 * s2;
  float rp = rs * (t3 - t4) / (t3 + t4);
  return (rs + rp) / 2.0f;
}
inline float ggx_d(float alpha, V3 wm) {  // ggx.rs:87-97
  float s0 = wm.x / alpha, s1 = wm.y / alpha;
  float t = wm.z * wm.z + s0 * s0 + s1 * s1;
  float a2 = alpha * alpha, t2 = t * t;
  return 1.0f / (PI_F * (a2 * t2));
}
inline float ggx_lambda(float alpha, V3 w) {  // ggx.rs:99-107
  if (w.z == 0.0f) return 0.0f;
  float a2 = alpha * alpha;
  float c = 1.0f + (a2 * (w.x * w.x) + a2 * (w.y * w.y)) / (w.z * w.z);
  return std::sqrt(c) * 0.5f - 0.5f;
}
inline float ggx_g(float alpha, V3 wi, V3 wo) { return 1.0f / (1.0f + ggx_lambda(alpha, wi) + ggx_lambda(alpha, wo)); }
inline float ggx_vnpdf(float alpha, V3 wi, V3 wh) {  // ggx.rs:115-119
  float inv_gl = 1.0f + ggx_lambda(alpha, wi);
  return (ggx_d(alpha, wh) * std::fabs(dot(wi, wh))) / (inv_gl * std::fabs(wi.z));
}
inline float ggx_vnpdf_no_d(float alpha, V3 wi, V3 wh) {  // ggx.rs:121-123
  return std::fabs(dot(wi, wh) / ((1.0f + ggx_lambda(alpha, wi)) * wi.z));
}
inline V3 sample_vndf(float alpha, V3 wi, float x, float y) {  // ggx.rs:129-169
  V3 v = normalized(v3(alpha * wi.x, alpha * wi.y, wi.z));
  V3 t1 = v.z < 0.9999f ? normalized(cross(v, v3(0, 0, 1))) : v3(1, 0, 0);
  V3 t2 = cross(t1, v);
  float a = 1.0f / (1.0f + v.z);
  float r = std::sqrt(x);
  float phi = y < a ? y / a * PI_F : PI_F + (y - a) / (1.0f - a) * PI_F;
  float sin_phi = std::sin(phi), cos_phi = std::cos(phi);
  float p1 = r * cos_phi;
  float p2 = r * sin_phi * (y < a ? 1.0f : v.z);
  float value = 1.0f - p1 * p1 - p2 * p2;
  V3 n = p1 * t1 + p2 * t2 + std::sqrt(std::fmax(value, 0.0f)) * v;
  return normalized(v3(alpha * n.x, alpha * n.y, std::fmax(n.z, 0.0f)));
}
inline V3 sample_wh(float alpha, V3 wi, float x, float y) {  // ggx.rs:171-180
  bool flip = wi.z < 0.0f;
  V3 wh = sample_vndf(alpha, flip ? -wi : wi, x, y);
  return flip ? -wh : wh;
}

struct GgxParams {
  float alpha, eta_inner, eta_outer, kappa;
  bool metallic;
};
inline float ggx_reflectance(const GgxParams &g, float cos_theta_i) {  // ggx.rs:221-227
  return g.metallic ? fresnel_conductor(g.eta_outer, g.eta_inner, g.kappa, cos_theta_i)
                    : fresnel_dielectric(g.eta_outer, g.eta_inner, cos_theta_i);
}
inline float ggx_reflectance_probability(const GgxParams &g, float cos_theta_i) {  // ggx.rs:229-242
  return g.metallic ? 1.0f : clampf(ggx_reflectance(g, cos_theta_i), 0.0f, 1.0f);
}
inline float ggx_eta_rel(const GgxParams &g, V3 wi) {  // ggx.rs:243-252
  return wi.z < 0.0f ? g.eta_outer / g.eta_inner : g.eta_inner / g.eta_outer;
}
// shared tail of bsdf() / generate_and_evaluate(): ggx.rs:462-473 / 286-303 (reflection lobe)
inline void ggx_reflect_terms(const GgxParams &g, V3 wi, V3 wo, V3 wh, float gcos, float ndotv, float &glossy, float &glossy_pdf) {
  float refl = ggx_reflectance(g, ndotv);
  float d = ggx_d(g.alpha, wh);
  float gg = ggx_g(g.alpha, wi, wo);
  glossy = refl * (0.25f / gcos) * d * gg;
  glossy_pdf = std::fabs(ndotv) == 0.0f ? 0.0f : ggx_vnpdf(g.alpha, wi, wh) * 0.25f / std::fabs(ndotv);
}
// ggx.rs:488-537 / 318-368 (transmission lobe), TransportMode::Importance
inline void ggx_transmit_terms(const GgxParams &g, V3 wi, V3 wo, V3 wh, float gcos, float &transmission, float &transmission_pdf) {
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
  transmission_pdf = std::fabs(d * partial * dwh_dwo2);
  float inv_reflectance = 1.0f - ggx_reflectance(g, ndotv);
  transmission = g.metallic ? 0.0f : inv_reflectance * std::fabs(weight);
}
// GGX::bsdf (ggx.rs:256-400)
inline Bsdf ggx_bsdf(const GgxParams &g, V3 wi, V3 wo) {
  wi = normalized(wi);
  bool same_hemisphere = wi.z * wo.z > 0.0f;
  float gcos = std::fabs(wi.z * wo.z);
  if (gcos == 0.0f) return {0.0f, 0.0f};
  float cos_i = wi.z;
  float glossy = 0.0f, transmission = 0.0f, glossy_pdf = 0.0f, transmission_pdf = 0.0f;
  if (same_hemisphere) {
    V3 wh = normalized(wo + wi);
    if (wh.z < 0.0f) wh = -wh;
    ggx_reflect_terms(g, wi, wo, wh, gcos, dot(wi, wh), glossy, glossy_pdf);
  } else if (!g.metallic) {
    float eta_rel = ggx_eta_rel(g, wi);
    V3 wh = normalized(wi + eta_rel * wo);
    if (wh.z < 0.0f) wh = -wh;
    ggx_transmit_terms(g, wi, wo, wh, gcos, transmission, transmission_pdf);
  }
  float refl_prob = ggx_reflectance_probability(g, cos_i);  // evaluated at wi.z (Q8)
  return {glossy + transmission, refl_prob * glossy_pdf + (1.0f - refl_prob) * transmission_pdf};
}
// GGX::generate_and_evaluate (ggx.rs:401-590)
inline Bsdf ggx_generate_and_evaluate(const GgxParams &g, float sx, float sy, V3 wi, V3 &wo) {
  V3 wh = normalized(sample_wh(g.alpha, wi, sx, sy));
  float refl_prob = ggx_reflectance_probability(g, dot(wh, wi));
  bool did_reflect = false;
  if (sx <= refl_prob) {  // same sample.x that drove the VNDF radius (Q7)
    did_reflect = true;
    wo = reflect(wi, wh);
  } else {
    float eta_rel = 1.0f / ggx_eta_rel(g, wi);
    if (!refract(wi, wh, eta_rel, wo)) {
      did_reflect = true;
      wo = reflect(wi, wh);
    }
  }
  float gcos = std::fabs(wi.z * wo.z);
  if (gcos == 0.0f) return {0.0f, 0.0f};
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
  return {glossy + transmission, rp * glossy_pdf + (1.0f - rp) * transmission_pdf};
}

inline GgxParams ggx_params(const RptScene &S, const RptMaterial &m, float lambda) {
  GgxParams g;
  g.alpha = m.alpha;
  g.eta_inner = curve_eval(S, m.curve_a, lambda);
  g.eta_outer = curve_eval(S, m.curve_b, lambda);
  g.metallic = m.metallic != 0;
  g.kappa = g.metallic ? curve_eval(S, m.curve_c, lambda) : 0.0f;
  return g;
}

// MaterialEnum::bsdf
inline Bsdf material_bsdf(const RptScene &S, uint32_t mat, float lambda, float u, float v, V3 wi, V3 wo) {
  const RptMaterial &m = S.materials[RPT_MAT_INDEX(mat)];
  switch (m.type) {
    case RPT_MATERIAL_LAMBERTIAN:  // lambertian.rs:16-32
      if (wo.z * wi.z > 0.0f) return {std::fmin(texstack_eval(S, m.texstack, lambda, u, v), 1.0f) / PI_F, std::fabs(wo.z) / PI_F};
      return {0.0f, 0.0f};
    case RPT_MATERIAL_GGX: return ggx_bsdf(ggx_params(S, m, lambda), wi, wo);
    default:  // diffuse_light.rs:29-45, sharp_light.rs:43-60
      if (wo.z * wi.z > 0.0f) return {clampf(curve_eval(S, m.curve_a, lambda), 0.0f, 1.0f) / PI_F, std::fabs(wo.z) / PI_F};
      return {0.0f, 0.0f};
  }
}
// MaterialEnum::generate_and_evaluate
inline Bsdf material_generate_and_evaluate(const RptScene &S, uint32_t mat, float lambda, float u, float v, float sx, float sy, V3 wi, V3 &wo) {
  const RptMaterial &m = S.materials[RPT_MAT_INDEX(mat)];
  switch (m.type) {
    case RPT_MATERIAL_LAMBERTIAN: {  // lambertian.rs:50-66
      wo = random_cosine_direction(sx, sy) * signum(wi.z);
      return {std::fmin(texstack_eval(S, m.texstack, lambda, u, v), 1.0f) / PI_F, std::fabs(wo.z) / PI_F};
    }
    case RPT_MATERIAL_GGX: return ggx_generate_and_evaluate(ggx_params(S, m, lambda), sx, sy, wi, wo);
    default: {  // diffuse_light.rs:60-76, sharp_light.rs:183-198
      wo = random_cosine_direction(sx, sy) * signum(wi.z);
      return {clampf(curve_eval(S, m.curve_a, lambda), 0.0f, 1.0f) / PI_F, std::fabs(wo.z) / PI_F};
    }
  }
}
// MaterialEnum::emission (diffuse_light.rs:123-133, sharp_light.rs:138-150,202-204; others 0)
inline float material_emission(const RptScene &S, uint32_t mat, float lambda, V3 wi) {
  const RptMaterial &m = S.materials[RPT_MAT_INDEX(mat)];
  if (m.type != RPT_MATERIAL_DIFFUSE_LIGHT && m.type != RPT_MATERIAL_SHARP_LIGHT) return 0.0f;
  float cosine = wi.z;
  bool ok = (cosine > 0.0f && m.sidedness == RPT_SIDED_FORWARD) || (cosine < 0.0f && m.sidedness == RPT_SIDED_REVERSE) ||
            m.sidedness == RPT_SIDED_DUAL;
  if (!ok) return 0.0f;
  float e = curve_eval(S, m.curve_b, lambda);
  if (m.type == RPT_MATERIAL_DIFFUSE_LIGHT) return e / PI_F;
  return e * ((m.sharpness + 1.0f) * std::pow(std::fabs(wi.z), m.sharpness) / 2.0f / PI_F);
}

// ---- curves with CDF (math::CurveWithCDF, Linear variant) -------------------------
enum InterpMode { MODE_LINEAR = 0, MODE_NEAREST = 1, MODE_CUBIC = 2 };
inline float interp(float t, float left, float right, int mode) {
  switch (mode) {
    case MODE_LINEAR: return (1.0f - t) * left + t * right;
    case MODE_NEAREST: return t < 0.5f ? left : right;
    default: {
      float t2 = 2.0f * t, omt = 1.0f - t;
      float h00 = (1.0f + t2) * omt * omt, h01 = t * t * (3.0f - t2);
      return h00 * left + h01 * right;
    }
  }
}
// Curve::Linear evaluate over `bounds`
inline float linear_curve_eval(const float *signal, uint32_t n, float lo, float hi, int mode, float x) {
  if (x < lo || x > hi) return 0.0f;
  float step = (hi - lo) / (float)n;
  uint32_t index = (uint32_t)((x - lo) / step);
  if (index >= n) index = n - 1;
  float left = signal[index];
  if (index + 1 >= n) return left;
  float right = signal[index + 1];
  float t = (x - (lo + (float)index * step)) / step;
  return interp(t, left, right, mode);
}
// CurveWithCDF::sample_power_and_pdf for a Linear cdf; returns abscissa and pdf = pdf(x)/pdf_integral.
inline void cdf_sample(const float *pdf, const float *cdf, uint32_t n, float lo, float hi, int mode, float pdf_integral,
                       float sample, float &x_out, float &pdf_out) {
  float lower_cdf = linear_curve_eval(cdf, n, lo, hi, mode, lo - 0.0001f);
  float upper_cdf = linear_curve_eval(cdf, n, lo, hi, mode, hi - 0.0001f);
  float s = lower_cdf + sample * (upper_cdf - lower_cdf);
  // binary_search_by_key: first index with cdf[i] >= s
  uint32_t index = (uint32_t)(std::lower_bound(cdf, cdf + n, s) - cdf);
  float x;
  if (index == 0) {
    x = lo;
  } else {
    if (index >= n) index = n - 1;
    float left = lo + ((float)index - 1.0f) * (hi - lo) / (float)n;
    float right = lo + (float)index * (hi - lo) / (float)n;
    float v0 = cdf[index - 1], v1 = cdf[index];
    float t = (s - v0) / (v1 - v0);
    x = interp(t, left, right, mode);
  }
  x_out = x;
  pdf_out = linear_curve_eval(pdf, n, lo, hi, mode, x) / pdf_integral;
}

// ---- environment (world/environment.rs) ------------------------------------------------------------------
inline float env_emission(const RptScene &S, float u, float v, float lambda) {  // :56-98
  const RptEnvironment &E = S.env;
  switch (E.kind) {
    case RPT_ENV_CONSTANT: return curve_eval(S, E.curve, lambda) * E.strength;
    case RPT_ENV_SUN: {
      V3 dir = uv_to_direction(u, v);
      float c = dot(v3(E.sun_direction[0], E.sun_direction[1], E.sun_direction[2]), dir);
      float s = std::sqrt(1.0f - c * c);
      if (std::fabs(s) < std::sin(E.angular_diameter / 2.0f) && c > 0.0f) return curve_eval(S, E.curve, lambda) * E.strength;
      return 0.0f;
    }
    default: {
      V3 dir = uv_to_direction_cr(u, v);
      V3 nd = mul_vec(*reinterpret_cast<const Mat4 *>(E.rot_reverse), dir);
      float uu, vv;
      direction_to_uv_cr(nd, uu, vv);
      return texstack_eval(S, E.texstack, lambda, uu, vv) * E.strength;
    }
  }
}
inline float env_pdf_for(const RptScene &S, float u, float v) {  // :198-258
  const RptEnvironment &E = S.env;
  switch (E.kind) {
    case RPT_ENV_CONSTANT: return 1.0f / (4.0f * PI_F);
    case RPT_ENV_SUN: {
      V3 dir = uv_to_direction(u, v);
      float c = dot(v3(E.sun_direction[0], E.sun_direction[1], E.sun_direction[2]), dir);
      float s = std::sqrt(1.0f - c * c);
      if (std::fabs(s) < std::sin(E.angular_diameter / 2.0f) && c > 0.0f) return 1.0f / (2.0f * PI_F * (1.0f - std::cos(E.angular_diameter)));
      return 0.0f;
    }
    default: {
      if (E.imap_rows == 0) return 1.0f / (4.0f * PI_F);
      V3 dir = uv_to_direction_cr(u, v);
      V3 nd = mul_vec(*reinterpret_cast<const Mat4 *>(E.rot_reverse), dir);
      float uu, vv;
      direction_to_uv_cr(nd, uu, vv);
      float m = linear_curve_eval(S.imap_m_pdf.data(), E.imap_marginal_n, 0.0f, 1.0f, MODE_NEAREST, uu);
      uint32_t row = (uint32_t)(clampf(uu, 0.0f, 1.0f - EPS_F) * (float)E.imap_rows);
      float r = linear_curve_eval(&S.imap_row_pdf[(size_t)row * E.imap_cols], E.imap_cols, 0.0f, 1.0f, MODE_NEAREST, vv);
      return m * r * (2.0f * PI_F * PI_F * std::sin(PI_F * vv) + 0.001f) + 0.001f;
    }
  }
}
inline void env_sample_uv(const RptScene &S, float sx, float sy, float &u, float &v, float &pdf) {  // :303-353
  const RptEnvironment &E = S.env;
  switch (E.kind) {
    case RPT_ENV_CONSTANT:
      u = sx;
      v = sy;
      pdf = 1.0f / (4.0f * PI_F);
      return;
    case RPT_ENV_SUN: {
      V3 local_wo = v3(0, 0, 1) + std::sin(E.angular_diameter / 2.0f) * random_in_unit_disk(sx, sy);
      Frame f = frame_from_normal(v3(E.sun_direction[0], E.sun_direction[1], E.sun_direction[2]));
      V3 dir = to_world(f, local_wo);
      direction_to_uv(normalized(dir), u, v);
      pdf = 1.0f / (2.0f * PI_F * (1.0f - std::cos(E.angular_diameter)));
      return;
    }
    default: {
      if (E.imap_rows == 0) {
        u = sx;
        v = sy;
        pdf = 1.0f / (4.0f * PI_F);
        return;
      }
      // ImportanceMap::sample_uv (importance_map.rs:325-357): sample.y -> row (u), sample.x -> column (v)
      float uu, row_pdf, vv, col_pdf;
      cdf_sample(S.imap_m_pdf.data(), S.imap_m_cdf.data(), E.imap_marginal_n, 0.0f, 1.0f, MODE_NEAREST, E.imap_marginal_integral, sy, uu, row_pdf);
      uint32_t row = (uint32_t)(uu * (float)E.imap_rows);
      if (row >= E.imap_rows) row = E.imap_rows - 1;
      cdf_sample(&S.imap_row_pdf[(size_t)row * E.imap_cols], &S.imap_row_cdf[(size_t)row * E.imap_cols], E.imap_cols, 0.0f, 1.0f, MODE_NEAREST, 1.0f, sx, vv, col_pdf);
      V3 local_wo = uv_to_direction_cr(uu, vv);
      V3 nw = mul_vec(*reinterpret_cast<const Mat4 *>(E.rot_forward), local_wo);
      direction_to_uv_cr(nw, u, v);
      pdf = row_pdf * col_pdf * (2.0f * PI_F * PI_F * std::sin(PI_F * v) + 0.001f) + 0.001f;
      return;
    }
  }
}

// ---- camera (camera/projective_camera.rs:101-120) -----------------------------------------------------
inline V3 a3(const float *p) { return v3(p[0], p[1], p[2]); }
inline Ray camera_get_ray(const RptCamera &C, float lens_x, float lens_y, float s, float t) {
  if (C.kind == RPT_CAMERA_PANORAMA) {  // camera/panorama_camera.rs:68-91; `transform.to_world(vec)` = u*x + v*y + w*z (math crate, unpinned)
    float ax = C.angle_span[0] * (s - 0.5f), ay = C.angle_span[1] * (0.5f - t);
    float sx = std::sin(ax), cx = std::cos(ax), sy = std::sin(ay), cy = std::cos(ay);
    V3 vec = v3(sx * cy, sy, cx * cy);
    return Ray{a3(C.origin), a3(C.u) * vec.x + a3(C.v) * vec.y + a3(C.w) * vec.z, INF_F};
  }
  V3 vec = random_in_unit_disk(lens_x, lens_y);  // optics::CircularAperture::sample (unpinned)
  V3 rd = C.aperture_diameter * vec;
  V3 offset = a3(C.u) * rd.x + a3(C.v) * rd.y;
  V3 origin = a3(C.origin) + offset;
  V3 pop = a3(C.lower_left) + s * a3(C.horizontal) + t * a3(C.vertical);
  return Ray{origin, normalized(pop - origin), INF_F};
}

// ---- the integrator -----------------------------------------------------------------------------------------
enum VType { VT_CAMERA, VT_EYE, VT_LIGHT_INSTANCE, VT_LIGHT_ENV };
struct Vertex {  // SurfaceVertex (integrator/utils.rs:39-55), fields the PT path reads
  VType type;
  V3 local_wi, point, normal;
  float u, v;
  uint32_t material, instance;
  float throughput, pdf_forward;
};

struct SampleCtx {
  uint64_t seed;
  uint32_t pixel, sample, light_samples;
};

struct Counters {
  uint64_t camera_rays = 0, bounce_rays = 0, shadow_rays = 0, env_hits = 0, segments = 0, true_rays = 0;
};

// random_walk (integrator/utils.rs:152-376), TransportMode::Importance, ignore_backward = true
inline void random_walk(const RptScene &S, Ray ray, float lambda, uint32_t bounce_limit, float start_throughput,
                        const SampleCtx &ctx, std::vector<Vertex> &vertices, uint32_t rr_start, Counters &cnt) {
  float beta = start_throughput;
  for (uint32_t bounce = 0; bounce < bounce_limit; ++bounce) {
    Hit hit;
    cnt.segments++;
    cnt.true_rays++;
    if (world_hit(S, ray, 0.0f, ray.tmax, hit)) {
      Frame frame = frame_from_normal(hit.n);
      V3 wi = normalized(to_local(frame, -ray.d));
      Vertex vx;
      vx.type = RPT_MAT_IS_LIGHT(hit.material) ? VT_LIGHT_INSTANCE : VT_EYE;
      vx.local_wi = wi;
      vx.point = hit.p;
      vx.normal = hit.n;
      vx.u = hit.u;
      vx.v = hit.v;
      vx.material = hit.material;
      vx.instance = hit.instance;
      vx.throughput = beta;
      vx.pdf_forward = 1.0f;
      RptRand4 s = rpt_philox(ctx.seed, ctx.pixel, ctx.sample, rpt_block_bsdf(bounce, ctx.light_samples));
      V3 wo;
      Bsdf fe = material_generate_and_evaluate(S, hit.material, lambda, hit.u, hit.v, s.x, s.y, wi, wo);
      float f = fe.f, pdf = fe.pdf;
      if (g_debug)
        std::printf("[walk b%u] p=(%.7g %.7g %.7g) n=(%.7g %.7g %.7g) mat=%u beta=%.7g f=%.7g pdf=%.7g wo=(%.7g %.7g %.7g) wi=(%.7g %.7g %.7g) s=(%.7g %.7g %.7g) inst=%u prim=%u t=%.7g\n",
                    bounce, hit.p.x, hit.p.y, hit.p.z, hit.n.x, hit.n.y, hit.n.z, hit.material, beta, f, pdf, wo.x, wo.y, wo.z, wi.x, wi.y, wi.z, s.x, s.y, s.z, hit.instance, hit.prim, hit.t);
      float cos_o = std::fabs(wo.z);
      if (std::isnan(pdf)) break;
      float rr = bounce >= rr_start ? std::fmin(f / pdf, 1.0f) : 1.0f;  // f32::min: NaN -> 1.0
      if (std::isnan(f / pdf) && bounce >= rr_start) rr = 1.0f;
      vx.pdf_forward = pdf * (rr / cos_o);
      vertices.push_back(vx);
      beta *= f / vx.pdf_forward;
      if (vx.pdf_forward == 0.0f) beta = 0.0f;
      if (beta == 0.0f) break;
      if (s.z > rr) break;
      ray = Ray{hit.p + hit.n * NORMAL_OFFSET * signum(wo.z), normalized(to_world(frame, wo)), INF_F};
    } else {
      Vertex vx;
      vx.type = VT_LIGHT_ENV;
      vx.local_wi = v3(0, 0, 1);
      vx.point = ray.d * S.world_radius;
      vx.normal = ray.d;
      vx.u = vx.v = 0.0f;
      vx.material = RPT_MAT_PACK(RPT_MAT_TAG_LIGHT, 0);
      vx.instance = 0;
      vx.throughput = beta;
      vx.pdf_forward = 0.0f;
      vertices.push_back(vx);
      break;
    }
  }
  cnt.bounce_rays += vertices.size();
}

// estimate_direct_illumination, live branch (pt.rs:146-219)
inline float estimate_direct_illumination(const RptScene &S, float lambda, const Vertex &vx, const Frame &frame, V3 wi,
                                          float throughput, bool only_direct, float pick, float sx, float sy, Counters &cnt) {
  size_t n = S.lights.size();
  if (n == 0) return 0.0f;
  size_t idx = (size_t)clampf((float)n * pick, 0.0f, (float)n - 1.0f);  // world/mod.rs:109
  uint32_t light = S.lights[idx];
  float pick_pdf = 1.0f / (float)n;
  V3 light_dir;
  float light_pdf;
  if (!instance_sample(S, light, sx, sy, vx.point, light_dir, light_pdf)) return 0.0f;
  light_pdf = light_pdf * pick_pdf;
  if (light_pdf == 0.0f) return 0.0f;
  V3 bsdf_wo = to_local(frame, light_dir);
  Bsdf b = material_bsdf(S, vx.material, lambda, vx.u, vx.v, wi, bsdf_wo);
  float weight = only_direct ? 1.0f : power_heuristic_generic(light_pdf, b.pdf);
  Ray shadow{vx.point + vx.normal * NORMAL_OFFSET * signum(bsdf_wo.z), light_dir, INF_F};
  cnt.shadow_rays++;
  cnt.true_rays++;
  Hit sh;
  if (g_debug)
    std::printf("[nee] dir=(%.7g %.7g %.7g) light_pdf=%.7g bsdf f=%.7g pdf=%.7g weight=%.7g bsdf_wo.z=%.7g\n", light_dir.x, light_dir.y, light_dir.z, light_pdf, b.f, b.pdf, weight, bsdf_wo.z);
  if (world_hit(S, shadow, 0.0f, INF_F, sh)) {
    if (RPT_MAT_IS_LIGHT(sh.material)) {
      Frame lf = frame_from_normal(sh.n);
      V3 lwi = to_local(lf, -light_dir);
      float le = material_emission(S, sh.material, lambda, lwi);
      float cos_i = std::fabs(lwi.z), cos_o = std::fabs(bsdf_wo.z);
      if (g_debug) std::printf("[shadow] hit inst=%u prim=%u t=%.7g le=%.7g lwi.z=%.7g -> %.7g\n", sh.instance, sh.prim, sh.t, le, lwi.z, b.f * throughput * cos_i * cos_o * le * weight / light_pdf);
      return b.f * throughput * cos_i * cos_o * le * weight / light_pdf;
    }
  }
  return 0.0f;
}
// estimate_direct_illumination_from_world (pt.rs:224-331)
inline float estimate_direct_illumination_from_world(const RptScene &S, float lambda, const Vertex &vx, const Frame &frame, V3 wi,
                                                     float throughput, bool only_direct, float sx, float sy, Counters &cnt) {
  float u, v, light_pdf;
  env_sample_uv(S, sx, sy, u, v, light_pdf);
  V3 direction = uv_to_direction(u, v);
  V3 local_wo = to_local(frame, direction);
  if (local_wo.z <= 0.0f) return 0.0f;
  Bsdf b = material_bsdf(S, vx.material, lambda, vx.u, vx.v, wi, local_wo);
  cnt.shadow_rays++;
  cnt.true_rays++;
  Hit sh;
  Ray shadow{vx.point + vx.normal * NORMAL_OFFSET * signum(direction.z), direction, INF_F};  // world z (Q12)
  if (world_hit(S, shadow, 0.0f, INF_F, sh)) return 0.0f;
  float emission = env_emission(S, u, v, lambda);
  float weight = only_direct ? 1.0f : power_heuristic_generic(light_pdf, b.pdf);
  return throughput * weight * b.f * emission * std::fabs(local_wo.z) * (1.0f / light_pdf);
}

// PathTracingIntegrator::color (pt.rs:397-615) -> energy for (pixel, sample); lambda returned.
inline float pt_color(const RptScene &S, const RptRenderParams &P, uint32_t px, uint32_t py, uint32_t sample, float &lambda_out,
                      Counters &cnt, Hit *primary = nullptr, bool *primary_hit = nullptr) {
  SampleCtx ctx{P.seed, py * P.width + px, sample, P.light_samples};
  cnt.camera_rays++;
  RptRand4 s0 = rpt_philox(ctx.seed, ctx.pixel, ctx.sample, 0);
  RptRand4 s1 = rpt_philox(ctx.seed, ctx.pixel, ctx.sample, 1);
  float cu = ((float)px + s0.x) / (float)P.width, cv = ((float)py + s0.y) / (float)P.height;  // tiled.rs:372-375
  float lambda = P.lambda_lo + s0.z * (P.lambda_hi - P.lambda_lo);                               // pt.rs:406
  lambda_out = lambda;
  float fu = clampf(cu, 0.0f, 1.0f - EPS_F), fv = clampf(cv, 0.0f, 1.0f - EPS_F);
  const RptCamera &C = S.cameras[P.camera];
  Ray camera_ray = camera_get_ray(C, s1.x, s1.y, fu, fv);
  uint32_t max_bounces = P.only_direct ? 1u : P.max_bounces;

  std::vector<Vertex> path;
  path.reserve(max_bounces + 1);
  Vertex first;
  first.type = VT_CAMERA;
  first.local_wi = v3(0, 0, 0);
  first.point = camera_ray.o;
  first.normal = camera_ray.d;
  first.u = first.v = 0.0f;
  first.material = 0;
  first.instance = 0;
  first.throughput = 1.0f;
  first.pdf_forward = 100.0f;
  path.push_back(first);

  if (primary) {  // parity hook (a): just the first closest hit
    *primary_hit = world_hit(S, camera_ray, 0.0f, camera_ray.tmax, *primary);
    return 0.0f;
  }
  random_walk(S, camera_ray, lambda, max_bounces, 1.0f, ctx, path, P.min_bounces, cnt);

  float energy = 0.0f;
  float p_env = S.lights.empty() ? 1.0f : S.p_env;  // world/mod.rs:170-176
  for (size_t index = 1; index < path.size(); ++index) {
    const Vertex &prev = path[index - 1];
    const Vertex &vx = path[index];
    if (vx.type == VT_LIGHT_ENV) {  // pt.rs:487-511
      V3 wo = vx.normal;
      float u, v;
      direction_to_uv(wo, u, v);
      float emission = env_emission(S, u, v, lambda);
      float cos_i = std::fabs(dot(prev.normal, wo));
      float nee_psa_pdf = env_pdf_for(S, u, v) / std::fabs(cos_i);
      float bsdf_psa_pdf = prev.pdf_forward / std::fabs(cos_i);
      float weight = power_heuristic(bsdf_psa_pdf, nee_psa_pdf);
      cnt.env_hits++;
      energy += weight * vx.throughput * emission;
    } else if (vx.type == VT_LIGHT_INSTANCE) {  // pt.rs:512-561
      float emission = material_emission(S, vx.material, lambda, vx.local_wi);
      if (g_debug) std::printf("[emit v%zu] emission=%.7g\n", index, emission);
      if (emission > 0.0f) {
        if (P.light_samples == 0 || prev.type == VT_CAMERA) {
          energy += vx.throughput * emission;
        } else if (P.only_direct) {
        } else {
          V3 nee_direction = normalized(vx.point - prev.point);
          float hyp = instance_psa_pdf(S, vx.instance, dot(prev.normal, nee_direction), dot(vx.normal, nee_direction), prev.point, vx.point);
          float weight = power_heuristic(prev.pdf_forward, hyp);
          energy += weight * vx.throughput * emission;
        }
      }
    } else {  // ordinary surface vertex: NEE (pt.rs:562-604, 333-393)
      Frame frame = frame_from_normal(vx.normal);
      V3 dir_to_prev = normalized(prev.point - vx.point);
      V3 wi = to_local(frame, dir_to_prev);
      if (P.light_samples > 0 && !(S.lights.empty() && p_env == 0.0f)) {
        float light_contribution = 0.0f;
        uint32_t bounce = (uint32_t)index - 1;
        for (uint32_t k = 0; k < P.light_samples; ++k) {
          RptRand4 s = rpt_philox(ctx.seed, ctx.pixel, ctx.sample, rpt_block_nee(bounce, ctx.light_samples, k));
          float pick;
          bool sample_world = choose(s.x, p_env, pick);
          if (sample_world)
            light_contribution += estimate_direct_illumination_from_world(S, lambda, vx, frame, wi, vx.throughput, P.only_direct != 0, s.y, s.z, cnt);
          else
            light_contribution += estimate_direct_illumination(S, lambda, vx, frame, wi, vx.throughput, P.only_direct != 0, pick, s.y, s.z, cnt);
        }
        energy += light_contribution / (float)P.light_samples;
      }
    }
  }
  return energy;
}

inline float cie_eval(const RptScene &S, int c, float lambda) {
  return lut_eval(&S.cie_lut[(size_t)c * S.num_lambda], S.num_lambda, S.lut_lo, S.lut_hi, lambda);
}

int fail(const std::string &msg) {
  g_error = msg;
  return 1;
}

}  // namespace

// ================================== exported C ABI (rpto_*) ===================================================
extern "C" {

const char *rpto_last_error(void) { return g_error.c_str(); }
uint32_t rpto_abi_version(void) { return RPT_ABI_VERSION; }

int rpto_scene_create(const RptSceneDesc *d, int /*device*/, RptScene **out) {
  if (!d || !out) return fail("null argument");
  if (d->abi_version != RPT_ABI_VERSION) return fail("ABI version mismatch");
  RptScene *S = new RptScene();
  S->instances.assign(d->instances, d->instances + d->num_instances);
  S->meshes.resize(d->num_meshes);
  for (uint32_t m = 0; m < d->num_meshes; ++m) {
    const RptMesh &src = d->meshes[m];
    OMesh &M = S->meshes[m];
    M.verts.resize(src.num_vertices);
    for (uint32_t i = 0; i < src.num_vertices; ++i) M.verts[i] = v3(src.vertices[3 * i], src.vertices[3 * i + 1], src.vertices[3 * i + 2]);
    if (src.normals) {
      M.normals.resize(src.num_vertices);
      for (uint32_t i = 0; i < src.num_vertices; ++i) M.normals[i] = v3(src.normals[3 * i], src.normals[3 * i + 1], src.normals[3 * i + 2]);
    }
    M.idx.assign(src.indices, src.indices + 3 * (size_t)src.num_faces);
    if (src.face_material) M.fmat.assign(src.face_material, src.face_material + src.num_faces);
    M.bbox = aabb_empty();
    for (auto &p : M.verts) M.bbox = aabb_grow(M.bbox, p);  // mesh.rs:271-274
    M.tri_aabb.resize(src.num_faces);
    for (uint32_t t = 0; t < src.num_faces; ++t)  // mesh.rs:57-64
      M.tri_aabb[t] = aabb_grow(aabb_new(M.verts[M.idx[3 * t]], M.verts[M.idx[3 * t + 1]]), M.verts[M.idx[3 * t + 2]]);
    M.bvh = build_flat_bvh(M.tri_aabb);
  }
  S->inst_aabb.resize(d->num_instances);
  AABB world_box = aabb_empty();
  for (uint32_t i = 0; i < d->num_instances; ++i) {
    const RptInstance &I = S->instances[i];
    AABB b;
    switch (I.kind) {
      case RPT_AGG_RECT: {  // rect.rs:58-66
        V3 v = shuffle_axis(v3(I.size[0] / 2.0f, I.size[1] / 2.0f, 0.0001f), I.axis);
        b = aabb_new(inst_origin(I) - v, inst_origin(I) + v);
        break;
      }
      case RPT_AGG_SPHERE: b = aabb_new(inst_origin(I) - v3(I.size[0], I.size[0], I.size[0]), inst_origin(I) + v3(I.size[0], I.size[0], I.size[0])); break;
      case RPT_AGG_DISK: {  // disk.rs:23-28 (half extent radius/2: reference quirk Q3)
        V3 v = v3(I.size[0] / 2.0f, I.size[0] / 2.0f, 0.001f);
        b = aabb_new(inst_origin(I) - v, inst_origin(I) + v);
        break;
      }
      default:
        if (I.mesh < 0 || (uint32_t)I.mesh >= d->num_meshes) {
          delete S;
          return fail("instance references a missing mesh");
        }
        b = S->meshes[I.mesh].bbox;
    }
    if (I.has_transform) b = aabb_transform(fwd(I), b);  // instance.rs:64-72
    S->inst_aabb[i] = b;
    world_box = aabb_expand(world_box, b);
  }
  S->tlas = build_flat_bvh(S->inst_aabb);
  S->world_radius = d->num_instances ? norm(world_box.mx - world_box.mn) / 2.0f : 0.0f;  // world/mod.rs:69-72
  S->lights.assign(d->lights, d->lights + d->num_lights);
  for (uint32_t l : S->lights)
    if (S->instances[l].kind == RPT_AGG_MESH) {
      delete S;
      return fail("mesh lights are unimplemented in the reference (src/geometry/mesh.rs:362-386 todo!())");
    }
  S->materials.assign(d->materials, d->materials + d->num_materials);
  S->num_lambda = d->num_lambda;
  S->lut_lo = d->lut_lambda_lo;
  S->lut_hi = d->lut_lambda_hi;
  S->curve_lut.assign(d->curve_lut, d->curve_lut + (size_t)d->num_curves * d->num_lambda);
  S->cie_lut.assign(d->cie_lut, d->cie_lut + 3 * (size_t)d->num_lambda);
  S->textures.resize(d->num_textures);
  for (uint32_t t = 0; t < d->num_textures; ++t) {
    const RptTexture &src = d->textures[t];
    OTexture &T = S->textures[t];
    T.channels = src.channels;
    T.w = src.width;
    T.h = src.height;
    T.texels.assign(src.texels, src.texels + (size_t)src.width * src.height * src.channels);
    std::memcpy(T.curves, src.curves, sizeof(T.curves));
  }
  S->stack_tex.assign(d->texstack_textures, d->texstack_textures + d->num_texstack_textures);
  S->stacks.assign(d->texstacks, d->texstacks + d->num_texstacks);
  S->env = d->environment;
  const RptEnvironment &E = d->environment;
  if (E.kind == RPT_ENV_HDR && E.imap_rows) {
    size_t n = (size_t)E.imap_rows * E.imap_cols;
    S->imap_row_pdf.assign(E.imap_row_pdf, E.imap_row_pdf + n);
    S->imap_row_cdf.assign(E.imap_row_cdf, E.imap_row_cdf + n);
    S->imap_m_pdf.assign(E.imap_marginal_pdf, E.imap_marginal_pdf + E.imap_marginal_n);
    S->imap_m_cdf.assign(E.imap_marginal_cdf, E.imap_marginal_cdf + E.imap_marginal_n);
  }
  S->p_env = d->num_lights == 0 ? 1.0f : d->env_sampling_probability;  // world/mod.rs:77-80
  S->cameras.assign(d->cameras, d->cameras + d->num_cameras);
  *out = S;
  return 0;
}

int rpto_scene_destroy(RptScene *S) {
  delete S;
  return 0;
}

// render_sampled (renderer/tiled.rs:279-542 / naive.rs:27-119): per pixel, per sample, box filter.
int rpto_render_pt(RptScene *S, const RptRenderParams *P, float *film, RptCounters *counters) {
  if (!S || !P || !film) return fail("null argument");
  if (P->camera >= S->cameras.size()) return fail("camera index out of range");
  const int64_t W = P->width, H = P->height;
  uint64_t c_cam = 0, c_bounce = 0, c_shadow = 0, c_env = 0, c_seg = 0, c_true = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : c_cam, c_bounce, c_shadow, c_env, c_seg, c_true)
  for (int64_t pix = 0; pix < W * H; ++pix) {
    uint32_t px = (uint32_t)(pix % W), py = (uint32_t)(pix / W);
    Counters cnt;
    float X = 0.0f, Y = 0.0f, Z = 0.0f;
    for (uint32_t s = 0; s < P->spp; ++s) {
      float lambda;
      float e = pt_color(*S, *P, px, py, P->spp_offset + s, lambda, cnt);
      X += e * cie_eval(*S, 0, lambda);  // XYZColor::from(SingleWavelength) (pt.rs:614)
      Y += e * cie_eval(*S, 1, lambda);
      Z += e * cie_eval(*S, 2, lambda);
    }
    float inv = P->spp_total ? 1.0f / (float)P->spp_total : 1.0f;  // tiled.rs:396-398
    film[4 * pix + 0] = X * inv;
    film[4 * pix + 1] = Y * inv;
    film[4 * pix + 2] = Z * inv;
    film[4 * pix + 3] = 0.0f;
    c_cam += cnt.camera_rays;
    c_bounce += cnt.bounce_rays;
    c_shadow += cnt.shadow_rays;
    c_env += cnt.env_hits;
    c_seg += cnt.segments;
    c_true += cnt.true_rays;
  }
  if (counters) {
    std::memset(counters, 0, sizeof(*counters));
    counters->camera_rays = c_cam;
    counters->bounce_rays = c_bounce;
    counters->shadow_rays = c_shadow;
    counters->env_hits = c_env;
    counters->segments = c_seg;
    counters->true_rays = c_true;
  }
  return 0;
}

// Debug: the walk vertices of one (pixel, sample). 16 floats per vertex:
// type, point xyz, normal xyz, throughput, pdf_forward, material, instance, u, v, 0,0,0. Returns the vertex count.
int rpto_debug_path(RptScene *S, const RptRenderParams *P, uint32_t px, uint32_t py, uint32_t sample, float *out, uint32_t cap) {
  SampleCtx ctx{P->seed, py * P->width + px, sample, P->light_samples};
  RptRand4 s0 = rpt_philox(ctx.seed, ctx.pixel, ctx.sample, 0);
  RptRand4 s1 = rpt_philox(ctx.seed, ctx.pixel, ctx.sample, 1);
  float cu = ((float)px + s0.x) / (float)P->width, cv = ((float)py + s0.y) / (float)P->height;
  float lambda = P->lambda_lo + s0.z * (P->lambda_hi - P->lambda_lo);
  Ray r = camera_get_ray(S->cameras[P->camera], s1.x, s1.y, clampf(cu, 0.0f, 1.0f - EPS_F), clampf(cv, 0.0f, 1.0f - EPS_F));
  std::vector<Vertex> path;
  Vertex first{};
  first.type = VT_CAMERA;
  first.point = r.o;
  first.normal = r.d;
  first.throughput = 1.0f;
  first.pdf_forward = 100.0f;
  path.push_back(first);
  Counters cnt;
  random_walk(*S, r, lambda, P->only_direct ? 1u : P->max_bounces, 1.0f, ctx, path, P->min_bounces, cnt);
  uint32_t n = 0;
  for (auto &v : path) {
    if (n >= cap) break;
    float *o = out + 16 * n++;
    o[0] = (float)v.type; o[1] = v.point.x; o[2] = v.point.y; o[3] = v.point.z;
    o[4] = v.normal.x; o[5] = v.normal.y; o[6] = v.normal.z; o[7] = v.throughput; o[8] = v.pdf_forward;
    o[9] = (float)RPT_MAT_INDEX(v.material); o[10] = (float)v.instance; o[11] = v.u; o[12] = v.v; o[13] = lambda; o[14] = o[15] = 0.0f;
  }
  return (int)n;
}

float rpto_debug_color(RptScene *S, const RptRenderParams *P, uint32_t px, uint32_t py, uint32_t sample) {
  g_debug = true;
  Counters cnt;
  float l;
  float e = pt_color(*S, *P, px, py, sample, l, cnt);
  g_debug = false;
  std::printf("[color] energy=%.7g lambda=%.7g\n", e, l);
  std::fflush(stdout);
  return e;
}

// Per-sample radiance dump for debugging / sample-level parity: energy[pix*spp + s], lambda likewise.
int rpto_render_samples(RptScene *S, const RptRenderParams *P, float *energy, float *lambda) {
  if (!S || !P || !energy) return fail("null argument");
  const int64_t W = P->width, H = P->height;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t pix = 0; pix < W * H; ++pix) {
    Counters cnt;
    for (uint32_t s = 0; s < P->spp; ++s) {
      float l;
      energy[pix * P->spp + s] = pt_color(*S, *P, (uint32_t)(pix % W), (uint32_t)(pix / W), P->spp_offset + s, l, cnt);
      if (lambda) lambda[pix * P->spp + s] = l;
    }
  }
  return 0;
}

int rpto_trace_primary(RptScene *S, const RptRenderParams *P, uint32_t *inst, uint32_t *prim, float *t) {
  if (!S || !P) return fail("null argument");
  const int64_t W = P->width, H = P->height;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t pix = 0; pix < W * H; ++pix) {
    Counters cnt;
    Hit h;
    bool hit = false;
    float l;
    pt_color(*S, *P, (uint32_t)(pix % W), (uint32_t)(pix / W), P->spp_offset, l, cnt, &h, &hit);
    inst[pix] = hit ? h.instance : NONE_U32;
    prim[pix] = hit ? h.prim : NONE_U32;
    t[pix] = hit ? h.t : INF_F;
  }
  return 0;
}

int rpto_trace_rays(RptScene *S, uint32_t n, const float *o, const float *d, const float *tmax, uint32_t *inst, uint32_t *prim, float *t) {
  if (!S) return fail("null argument");
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    Ray r{v3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), v3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), tmax[i]};
    Hit h;
    bool hit = world_hit(*S, r, 0.0f, r.tmax, h);
    inst[i] = hit ? h.instance : NONE_U32;
    prim[i] = hit ? h.prim : NONE_U32;
    t[i] = hit ? h.t : INF_F;
  }
  return 0;
}

// ---- output_film (renderer/mod.rs:24-80): Tonemapper::initialize + map (tonemap/{clamp,reinhard0,reinhard1}.rs),
// XYZ -> RGB (tonemap/mod.rs:24-40,116-140), OETF (:147-205), byte encoding (:314-331). Sequential like the reference.
namespace {
const float XYZ_TO_REC709[9] = {3.24096994f, -1.53738318f, -0.49861076f, -0.96924364f, 1.8759675f, 0.04155506f, 0.05563008f, -0.20397696f, 1.05697151f};
const float XYZ_TO_REC2020[9] = {1.4628067f, -0.1840623f, -0.2743606f, -0.5217933f, 1.4472381f, 0.0677227f, 0.0349342f, -0.0968930f, 1.2884099f};
const float MAUVE_XYZ[3] = {0.5199467f, 51.48687f, 1.0180528f};  // src/lib.rs:46
inline float oetf(float v, uint32_t cs) {
  if (cs == RPT_COLORSPACE_SRGB) return v < 0.0031308f ? (323.0f / 25.0f) * v : (211.0f / 200.0f) * std::pow(v, 5.0f / 12.0f) - (11.0f / 200.0f);
  return v < 0.01805397f ? 4.5f * v : 1.0992968f * std::pow(v, 0.45f) - 0.09929682f;
}
inline bool finite3(const float *c) { return std::isfinite(c[0]) && std::isfinite(c[1]) && std::isfinite(c[2]) && std::isfinite(c[3]); }
}  // namespace

int rpto_output_film(RptScene *, const float *film, uint32_t W, uint32_t H, const RptOutputSettings *O, float *rgb_linear, uint8_t *rgba8, float *l_w_out) {
  if (!film || !O || !rgba8) return fail("null argument");
  if (!(O->factor > 0.0f)) return fail("factor must be > 0 (renderer/mod.rs:26)");
  const size_t n = (size_t)W * H;
  const float *M = O->colorspace == RPT_COLORSPACE_REC2020 ? XYZ_TO_REC2020 : XYZ_TO_REC709;
  // ---- initialize
  float lw[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  const bool x3 = !O->luminance_only && O->tonemapper != RPT_TONEMAP_CLAMP;
  if (O->tonemapper != RPT_TONEMAP_CLAMP) {
    double sum_log = 0.0;
    float sum_log4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (size_t i = 0; i < n; ++i) {
      const float *c = film + 4 * i;
      float lum = c[1];
      if (std::isnan(lum)) continue;
      if (x3) {
        for (int k = 0; k < 4; ++k) sum_log4[k] += std::log(0.001f + c[k]);
      } else if (O->tonemapper == RPT_TONEMAP_REINHARD0) {
        sum_log += std::log(0.001 + (double)lum);  // reinhard0.rs:48 (f64 DELTA)
      } else {
        sum_log += std::log((double)(0.001f + lum));  // reinhard1.rs:47 (f32 add, then f64)
      }
    }
    if (x3)
      for (int k = 0; k < 4; ++k) lw[k] = std::exp(sum_log4[k] / (float)n) / O->factor;
    else
      lw[0] = lw[1] = lw[2] = lw[3] = (float)std::exp(sum_log / (double)n) / O->factor;
  }
  if (l_w_out) std::memcpy(l_w_out, lw, sizeof(lw));
  // ---- per pixel
  for (size_t i = 0; i < n; ++i) {
    const float *c = film + 4 * i;
    if (rgb_linear) {  // EXR payload: M * (factor * XYZ) (tonemap/mod.rs:225-247)
      float x = O->factor * c[0], y = O->factor * c[1], z = O->factor * c[2];
      for (int r = 0; r < 3; ++r) rgb_linear[3 * i + r] = M[3 * r] * x + M[3 * r + 1] * y + M[3 * r + 2] * z;
    }
    float m[3];
    if (O->tonemapper == RPT_TONEMAP_CLAMP) {  // clamp.rs:76-101
      float v[4] = {c[0] * O->factor, c[1] * O->factor, c[2] * O->factor, c[3] * O->factor};
      if (!finite3(v)) { v[0] = MAUVE_XYZ[0]; v[1] = MAUVE_XYZ[1]; v[2] = MAUVE_XYZ[2]; }
      float em = std::pow(2.0f, O->exposure);
      if (O->luminance_only) {
        float lum = v[1];
        float new_lum = clampf(lum * em, 0.0f, 1.0f);
        float sf = new_lum / lum;
        for (int k = 0; k < 3; ++k) m[k] = sf * v[k];
      } else {
        for (int k = 0; k < 3; ++k) m[k] = std::fmax(std::fmin(v[k] * em, 1.0f), 0.0f);
      }
    } else if (!x3) {  // reinhard0.rs:96-113 / reinhard1.rs:88-107
      float lum = c[1];
      float l = O->key_value * lum / lw[1];
      float sf;
      if (O->tonemapper == RPT_TONEMAP_REINHARD0) {
        sf = l / (1.0f + l);
      } else {
        float mul = 1.0f / (O->white_point * O->white_point);
        sf = l * (mul * l + 1.0f) / (1.0f + l);
      }
      float v[3] = {c[0], c[1], c[2]};
      if (!finite3(c)) { v[0] = MAUVE_XYZ[0]; v[1] = MAUVE_XYZ[1]; v[2] = MAUVE_XYZ[2]; }
      for (int k = 0; k < 3; ++k) m[k] = sf * v[k];
    } else {  // reinhard0.rs:196-213 / reinhard1.rs:198-232 (per channel)
      bool bad = false;
      float v[3] = {c[0], c[1], c[2]};
      if (O->tonemapper == RPT_TONEMAP_REINHARD0 && !finite3(c)) { v[0] = MAUVE_XYZ[0]; v[1] = MAUVE_XYZ[1]; v[2] = MAUVE_XYZ[2]; }
      for (int k = 0; k < 3; ++k) {
        float l = O->key_value * c[k] / lw[k];
        float sf;
        if (O->tonemapper == RPT_TONEMAP_REINHARD0) {
          sf = l / (1.0f + l);
        } else {
          float mul = 1.0f / std::pow(O->white_point, 2.0f);
          sf = l * (mul * l + 1.0f) / (1.0f + l);
        }
        m[k] = sf * v[k];
        if (!std::isfinite(m[k])) bad = true;
      }
      if (O->tonemapper == RPT_TONEMAP_REINHARD1 && bad) { m[0] = MAUVE_XYZ[0]; m[1] = MAUVE_XYZ[1]; m[2] = MAUVE_XYZ[2]; }
    }
    for (int r = 0; r < 3; ++r) {
      float lin = M[3 * r] * m[0] + M[3 * r + 1] * m[1] + M[3 * r + 2] * m[2];
      float e = std::ceil(oetf(lin, O->colorspace) * 255.0f);
      e = e < 0.0f ? 0.0f : (e > 255.0f ? 255.0f : e);  // NaN stays NaN -> `as u8` = 0
      rgba8[4 * i + r] = std::isnan(e) ? 0 : (uint8_t)e;
    }
    rgba8[4 * i + 3] = 255;
  }
  return 0;
}

// ---- N3: ImportanceMap::bake_raw (world/importance_map.rs:78-253), scalar restatement -----------------------------
// Curve::Machine (math crate, unpinned): seed, then every (op, curve) applied left to right, result clamped at 0.
// Curve::evaluate_integral(bounds, n, clamped = false): left Riemann sum, f32 accumulation, times the step (unpinned).
int rpto_scene_bake_importance_map(RptScene *S, const RptImapBake *B, float *row_pdf, float *row_cdf, float *marginal_pdf, float *marginal_cdf,
                                   float *marginal_integral) {
  if (!S || !B) return fail("null argument");
  if (S->env.kind != RPT_ENV_HDR) return fail("importance maps exist for HDR environments only");
  if (B->rows == 0 || B->cols == 0 || B->num_samples == 0) return fail("empty importance map");
  const uint32_t R = B->rows, Cn = B->cols, NS = B->num_samples;
  const RptTexStack &st = S->stacks[S->env.texstack];
  const float step = (B->lambda_hi - B->lambda_lo) / (float)NS;
  std::vector<float> lum((size_t)R * Cn), row_sum(R);
  S->imap_row_pdf.assign((size_t)R * Cn, 0.0f);
  S->imap_row_cdf.assign((size_t)R * Cn, 0.0f);
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t row = 0; row < (int64_t)R; ++row) {
    float row_luminance = 0.0f;
    for (uint32_t col = 0; col < Cn; ++col) {
      float u = (float)row / (float)R, v = (float)col / (float)Cn;  // :137-140
      float sum = 0.0f;
      for (uint32_t i = 0; i < NS; ++i) {
        float stack = 0.0f;  // TexStack::curve_at: Machine{0, [Add tex.curve_at(uv)]} (texture.rs:221-228)
        for (uint32_t k = 0; k < st.count; ++k) {
          const OTexture &T = S->textures[S->stack_tex[st.first + k]];
          float uu = clampf(u, 0.0f, 1.0f - EPS_F), vv = clampf(v, 0.0f, 1.0f - EPS_F);  // vec2d.rs:34-42
          size_t x = (size_t)(uu * (float)T.w), y = (size_t)(vv * (float)T.h);
          const float *tx = &T.texels[(y * T.w + x) * T.channels];
          const float *bs = B->basis + (size_t)(4 * k) * NS;
          float tv;
          if (T.channels == 1) {
            tv = std::fmax(tx[0] * bs[i], 0.0f);  // Texture1::curve_at: Machine{texel, [Mul curve]} (texture.rs:126-131)
          } else {
            tv = 0.0f;  // Texture4::curve_at: Machine{0, [Add Machine{texel[c], [Mul curves[c]]}]} (texture.rs:41-77)
            for (int c = 0; c < 4; ++c) tv = tv + std::fmax(tx[c] * bs[(size_t)c * NS + i], 0.0f);
            tv = std::fmax(tv, 0.0f);
          }
          stack = stack + tv;
        }
        stack = std::fmax(stack, 0.0f);
        sum += std::fmax(1.0f * B->luminance[i] * stack, 0.0f);  // Machine{1, [Mul luminance, Mul stack curve]} (:141-147)
      }
      float texel_luminance = sum * step;
      row_luminance += texel_luminance;  // :153
      lum[(size_t)row * Cn + col] = texel_luminance;
      S->imap_row_cdf[(size_t)row * Cn + col] = row_luminance;
    }
    for (uint32_t col = 0; col < Cn; ++col) {  // :158-163 (a black row divides 0 by 0, as the reference does)
      S->imap_row_pdf[(size_t)row * Cn + col] = lum[(size_t)row * Cn + col] / row_luminance;
      S->imap_row_cdf[(size_t)row * Cn + col] /= row_luminance;
    }
    row_sum[row] = row_luminance;
  }
  float total = 0.0f;
  for (uint32_t r = 0; r < R; ++r) total += row_sum[r];  // :199
  S->imap_m_pdf.resize(R);
  S->imap_m_cdf.resize(R);
  for (uint32_t r = 0; r < R; ++r) S->imap_m_pdf[r] = row_sum[r] / total;  // :214
  // Curve::Linear::to_cdf (math crate, unpinned): running sum of signal * step over the curve's own samples, normalised
  float mstep = (1.0f - 0.0f) / (float)R, acc = 0.0f;
  for (uint32_t r = 0; r < R; ++r) {
    acc += S->imap_m_pdf[r] * mstep;
    S->imap_m_cdf[r] = acc;
  }
  float integral = acc;
  for (uint32_t r = 0; r < R; ++r) S->imap_m_cdf[r] /= integral;
  S->env.imap_rows = R;
  S->env.imap_cols = Cn;
  S->env.imap_marginal_n = R;
  S->env.imap_marginal_integral = integral;
  size_t n = (size_t)R * Cn;
  if (row_pdf) std::copy(S->imap_row_pdf.begin(), S->imap_row_pdf.end(), row_pdf);
  if (row_cdf) std::copy(S->imap_row_cdf.begin(), S->imap_row_cdf.end(), row_cdf);
  if (marginal_pdf) std::copy(S->imap_m_pdf.begin(), S->imap_m_pdf.end(), marginal_pdf);
  if (marginal_cdf) std::copy(S->imap_m_cdf.begin(), S->imap_m_cdf.end(), marginal_cdf);
  if (marginal_integral) *marginal_integral = integral;
  (void)n;
  return 0;
}

// ---- unit hooks for the known-answer tests (tests/test_oracle_*.py) ----------------------------------------
// ggx_glass(roughness) of the reference's tests: eta = cauchy(1.5, 10000), eta_o = 1, kappa = 0 (ggx.rs:630-635)
void rpto_ggx_bsdf(float alpha, float eta_inner, float eta_outer, float kappa, int metallic, const float *wi, const float *wo, float *f, float *pdf) {
  GgxParams g{alpha, eta_inner, eta_outer, kappa, metallic != 0};
  Bsdf b = ggx_bsdf(g, a3(wi), a3(wo));
  *f = b.f;
  *pdf = b.pdf;
}
void rpto_ggx_generate_and_evaluate(float alpha, float eta_inner, float eta_outer, float kappa, int metallic, float sx, float sy,
                                    const float *wi, float *wo, float *f, float *pdf) {
  GgxParams g{alpha, eta_inner, eta_outer, kappa, metallic != 0};
  V3 o;
  Bsdf b = ggx_generate_and_evaluate(g, sx, sy, a3(wi), o);
  wo[0] = o.x;
  wo[1] = o.y;
  wo[2] = o.z;
  *f = b.f;
  *pdf = b.pdf;
}
float rpto_fresnel_dielectric(float eta_i, float eta_t, float cos_i) { return fresnel_dielectric(eta_i, eta_t, cos_i); }
float rpto_fresnel_conductor(float eta_i, float eta_t, float k, float cos_i) { return fresnel_conductor(eta_i, eta_t, k, cos_i); }
void rpto_cdf_sample(const float *pdf, const float *cdf, uint32_t n, float lo, float hi, int mode, float pdf_integral, float sample, float *x, float *p) {
  cdf_sample(pdf, cdf, n, lo, hi, mode, pdf_integral, sample, *x, *p);
}
float rpto_linear_curve_eval(const float *signal, uint32_t n, float lo, float hi, int mode, float x) { return linear_curve_eval(signal, n, lo, hi, mode, x); }
void rpto_frame_roundtrip(const float *n, const float *v, float *local, float *world) {
  Frame f = frame_from_normal(a3(n));
  V3 l = to_local(f, a3(v)), w = to_world(f, l);
  local[0] = l.x; local[1] = l.y; local[2] = l.z;
  world[0] = w.x; world[1] = w.y; world[2] = w.z;
}
void rpto_philox(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t block, float *out4) {
  RptRand4 r = rpt_philox(seed, pixel, sample, block);
  out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
// flat BVH of the TLAS for structure tests: returns node count; fills up to cap (entry, exit, shape) triples
uint32_t rpto_tlas_nodes(RptScene *S, uint32_t *triples, uint32_t cap) {
  uint32_t n = (uint32_t)S->tlas.size();
  for (uint32_t i = 0; i < n && i < cap; ++i) {
    triples[3 * i] = S->tlas[i].entry;
    triples[3 * i + 1] = S->tlas[i].exit;
    triples[3 * i + 2] = S->tlas[i].shape;
  }
  return n;
}
// torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm of the benchmark asks for the host's cores explicitly
void rpto_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int rpto_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
