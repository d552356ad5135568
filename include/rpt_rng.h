/*
 * rpt_rng.h — counter-based sample generator shared by the CUDA path and the CPU oracle.
 *
 * The reference draws from `RandomSampler` (thread-local OS-seeded `rand`, reference
 * src/renderer/tiled.rs:344, naive.rs:79), so it is not reproducible run to run. Both of our
 * implementations instead use Philox4x32-10 (Salmon et al., SC'11) keyed by (seed) and
 * counted by (pixel, global sample index, dimension block): the SAME inputs on both sides,
 * which is what makes sample-level GPU-vs-oracle comparison possible.
 *
 * Dimension blocks of one camera sample (draw order of reference pt.rs:397-615):
 *   block 0                      : x,y = film jitter (tiled.rs:369), z = wavelength (pt.rs:406)
 *   block 1                      : x,y = lens / aperture sample (projective_camera.rs:102)
 *   block 2 + b*(1+L)            : x,y = BSDF sample, z = russian roulette   (walk bounce b)
 *   block 2 + b*(1+L) + 1 + k    : x = light-vs-env choice / light pick, y,z = light sample
 *                                  (NEE sample k at the vertex found by bounce b), L = light_samples
 */
#ifndef RPT_RNG_H
#define RPT_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RPT_HD __host__ __device__ __forceinline__
#else
#define RPT_HD static inline
#endif

typedef struct RptRand4 {
  float x, y, z, w;
} RptRand4;

RPT_HD uint32_t rpt_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* uniform in [0,1): top 24 bits */
RPT_HD float rpt_u32_to_unit(uint32_t v) { return (float)(v >> 8) * (1.0f / 16777216.0f); }

RPT_HD RptRand4 rpt_philox(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t block) {
  uint32_t c0 = pixel, c1 = sample, c2 = block, c3 = 0x52505442u; /* "RPTB" */
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = rpt_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = rpt_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  RptRand4 o;
  o.x = rpt_u32_to_unit(c0);
  o.y = rpt_u32_to_unit(c1);
  o.z = rpt_u32_to_unit(c2);
  o.w = rpt_u32_to_unit(c3);
  return o;
}

RPT_HD uint32_t rpt_block_bsdf(uint32_t bounce, uint32_t light_samples) { return 2u + bounce * (1u + light_samples); }
RPT_HD uint32_t rpt_block_nee(uint32_t bounce, uint32_t light_samples, uint32_t k) {
  return 2u + bounce * (1u + light_samples) + 1u + k;
}

#endif /* RPT_RNG_H */
