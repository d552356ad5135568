// rpt_kernels.cu — the wavefront PT pipeline (sm_100a) and the C ABI of include/rpt.h.
//
// Pipeline per wave of N = width*height*spp_chunk camera samples (reference pt.rs:397-615 unrolled
// into a bounce-synchronous wavefront):
//   k_raygen            : film jitter, wavelength, camera ray -> path queue (fused into bounce 0 of k_trace in a render)
//   per bounce b:
//     k_trace           : two-level BVH closest hit, sorts paths by material class  -> hit records + class lists
//     k_shade_miss      : environment vertex: emission * MIS                         (pt.rs:487-511)
//     k_shade_surface<> : one launch per material class: light-hit MIS, NEE sample generation,
//                         BSDF sampling, russian roulette                            -> next path queue + shadow queue
//     k_shadow          : NEE visibility (closest hit for lights, any hit for env)  -> per-sample energy
//   k_film              : energy * (x_bar, y_bar, z_bar)(lambda) summed per pixel   -> XYZ film
// Queues are compacted with warp ballot / popc and one atomic per warp per queue; queue sizes stay
// on the device, every kernel is a persistent grid-stride launch, so a wave needs no host sync.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "rpt_bvh.h"
#include "rpt_device.cuh"

#ifdef RPT_DEBUG
// debug builds only (make debug): kernel printf for one slot, selected with rpt_debug_set_slot()
__device__ uint32_t d_debug_slot = 0xFFFFFFFFu;
#define RPT_DEBUG_SLOT d_debug_slot
extern "C" int rpt_debug_set_slot(uint32_t slot) { return cudaMemcpyToSymbol(d_debug_slot, &slot, sizeof(slot)) == cudaSuccess ? 0 : 1; }
#endif

namespace {

thread_local std::string g_error;
std::mutex g_cache_mu;  // guards the process-global per-device caches (DeviceBuffers::spare, g_wave_cache) and cudaFuncSetAttribute
int fail(const std::string &msg) {
  g_error = msg;
  return 1;
}
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)

// ---------------------------------------------------------------------------------------------
// queue records
// ---------------------------------------------------------------------------------------------
// Path record, 64 B = 4 x float4, written contiguously by the producer (coalesced 128-bit stores).
//   r0 = prev.point.xyz, beta                 (throughput before this segment's vertex)
//   r1 = prev.normal.xyz, prev.pdf_forward    (camera vertex: ray direction, 100; pt.rs:436,441)
//   r2 = ray direction.xyz, lambda
//   r3 = slot (bits), offset sign (0 for the camera ray, else signum(wo.z)), unused, unused
// The ray origin is derived exactly as the reference derives it: prev.point + prev.normal*0.001*sign
// (integrator/utils.rs:326-329).
struct __align__(16) PathRec {
  float4 r0, r1, r2, r3;
};
struct __align__(16) HitRec {
  float t;
  uint32_t inst, prim;
  uint32_t mat;  // packed MaterialId of the hit (what k_trace classified it by): lets the shade kernel start its material fetch at once
};
struct __align__(16) NeeRec {  // NEE hand-over record of the split shade pipeline (layout: see k_shade_vertex)
  float4 r0, r1, r2, r3;
};

// Per-bounce counters. Q_PATHS..Q_SHADOW are queue SIZES (they include the invalid entries that pad abandoned
// chunk tails, see chunk_append); the N_* entries count valid items and feed the Profile counters.
enum : uint32_t {
  Q_PATHS = 0, Q_MISS = 1, Q_DIFFUSE = 2, Q_GGX = 3, Q_SHADOW = 4, Q_NAN = 5, Q_SHADOW_REF = 6,
  N_PATHS = 7, N_MISS = 8, N_DIFFUSE = 9, N_GGX = 10, N_SHADOW = 11,
  F_TRACE = 12, F_SHADOW = 13, F_SHADE_DIFFUSE = 14, F_SHADE_GGX = 15,  // claim counters of the dynamic tile hand-out (TileStream)
  Q_NEE_DIFFUSE = 16, Q_NEE_GGX = 17,  // sizes of the two NEE hand-over queues (split shade pipeline), padding included
  N_NEE = 18,                          // valid NEE hand-over records (both classes)
  F_NEE_DIFFUSE = 19, F_NEE_GGX = 20, F_NEE_DIFFUSE_ENV = 21, F_NEE_GGX_ENV = 22,
  Q_COUNT = 24
};

struct RenderCtx {
  uint32_t width, height, wh;
  uint32_t tiles_x;        // 0: slot q of a frame is pixel q; else width / 8: slots run over 8 x 4 pixel tiles (slot_to_pixel)
  uint32_t tiles_magic;    // ceil(2^32 / tiles_x) when the multiply-high division by tiles_x is exact for every tile of the frame, else 0
  uint32_t n_slots;        // slots in this wave
  uint32_t sample_base;    // global sample index of slot 0's sample
  uint32_t min_bounces, max_bounces, light_samples, only_direct;
  float lambda_lo, lambda_hi;
  uint64_t seed;
  RptCamera cam;
};

struct WaveBuffers {
  PathRec *paths[2];
  HitRec *hits;
  uint32_t *q_miss, *q_diffuse, *q_ggx;
  NeeRec *nee_d, *nee_g;  // NEE hand-over queues, one per material class
  float4 *sh_a, *sh_b;  // shadow records: (origin.xyz, pre-contribution), (dir.xyz, lambda)
  uint32_t *sh_c;       // slot | kind << 31 (1 = environment any-hit)
  float *acc;           // per-slot energy (pt.rs `sum.energy`)
  uint32_t *counts;     // [bounces + 1][Q_COUNT] (RptScene::counts_cap rows)
  unsigned long long *work;  // [2][3]: (nodes, triangles, instances) visited by k_trace / k_shadow
};

// Which pixel the q-th slot of a frame belongs to. Row-major, a warp's 32 camera rays are a 32 x 1 strip of pixels; with
// tiles (film width a multiple of 8, height a multiple of 4) they are an 8 x 4 block, whose rays - and the paths that follow
// them through the queues - stay closer together in the scene. Samples are keyed by (pixel, sample index) and the film is
// accumulated per pixel, so the mapping changes the order of the work, not the image.
__device__ __forceinline__ uint32_t slot_to_pixel(const RenderCtx &R, uint32_t q) {
  if (R.tiles_x == 0u) return q;
  const uint32_t tile = q >> 5, l = q & 31u;
  const uint32_t ty = R.tiles_magic ? __umulhi(tile, R.tiles_magic) : tile / R.tiles_x, tx = tile - ty * R.tiles_x;
  return ((ty << 2) + (l >> 3)) * R.width + (tx << 3) + (l & 7u);
}

__device__ __forceinline__ float3 rec_origin(const PathRec &r) {
  float3 p = f3(r.r0), n = f3(r.r1);
  return p + (n * RPT_NORMAL_OFFSET) * r.r3.y;
}

// cp.async (LDGSTS) helpers: 16-byte global -> shared copies that bypass L1 (.cg) and complete asynchronously.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers: one lane stages a warp's whole queue tile
// (32 consecutive records) into shared memory while the warp is still tracing the previous tile.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RPT_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RPT_DONE_%=;\n"
      "bra RPT_WAIT_%=;\n"
      "RPT_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Queue append without a global atomic per warp per iteration. Every warp owns a private chunk of QCHUNK
// consecutive entries of the output queue and hands them out with ballot + popc; only when the chunk is used up
// does one lane reserve the next chunk with a single atomicAdd on the queue size (ncu on the first version: 46 % of
// the shade kernel's stall samples sat on the SHFL that waits for the per-warp atomic's return value).
// The tail a warp abandons (at most 31 entries when a chunk overflows, the rest of the chunk at kernel end) is
// filled with invalid markers (RPT_NONE in the entry's first id word); consumers skip them. Chunk tails are
// contiguous, so whole warps of the consumer skip together.
// Chunk size may follow the size of the kernel's INPUT queue (QCHUNK_SMALL / QCHUNK_BINNED_SMALL for launches of fewer
// than SMALL_QUEUE / BIN_MIN_ITEMS entries). With 256-entry chunks every one of the ~4 700 resident warps abandons on
// average half a chunk per queue per launch, and for the late bounces (0.5-3 M live paths) that padding is as large as the
// payload. Measured in one session (profiles/r02_chunk_policy.md): 32-entry chunks below 4 M entries help a 512x384 frame
// (7.12 -> 6.55 ms) but cost the 1080p Cornell frame 7 % (20.99 -> 22.43 ms; the vertex kernel 3.97 -> 4.77 ms) and the gem
// 3 %: four times the same-address atomicAdds on the queue counters outweigh the padding. The defaults therefore keep one
// chunk size; the policy stays compiled in for the record (-DQCHUNK_SMALL=32u -DQCHUNK_BINNED_SMALL=32u).
#define QCHUNK 256u
#ifndef QCHUNK_SMALL
#define QCHUNK_SMALL 256u
#endif
#ifndef QCHUNK_BINNED_SMALL
#define QCHUNK_BINNED_SMALL 128u
#endif
#ifndef SMALL_QUEUE
#define SMALL_QUEUE (4u << 20)
#endif
struct WarpChunk {
  uint32_t base, used;
};
__device__ __forceinline__ uint32_t chunk_size_for(uint32_t n_in) { return n_in >= SMALL_QUEUE ? QCHUNK : QCHUNK_SMALL; }
__device__ __forceinline__ WarpChunk chunk_init(uint32_t cap = QCHUNK) { return WarpChunk{0u, cap}; }
template <class Mark>
__device__ __forceinline__ void chunk_pad(const WarpChunk &wc, Mark mark, uint32_t cap = QCHUNK) {
  for (uint32_t e = wc.used + (threadIdx.x & 31u); e < cap; e += 32u) mark(wc.base + e);
}
// All 32 lanes must call. Returns this lane's entry index, or RPT_NONE if !pred.
template <class Mark>
__device__ __forceinline__ uint32_t chunk_append(uint32_t *counter, WarpChunk &wc, bool pred, Mark mark, uint32_t cap = QCHUNK) {
  uint32_t mask = __ballot_sync(0xFFFFFFFFu, pred);
  if (mask == 0) return RPT_NONE;
  uint32_t lane = threadIdx.x & 31u;
  uint32_t cnt = __popc(mask);
  if (wc.used + cnt > cap) {  // warp-uniform
    chunk_pad(wc, mark, cap);
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(counter, cap);
    wc.base = __shfl_sync(0xFFFFFFFFu, base, 0);
    wc.used = 0;
  }
  uint32_t idx = wc.base + wc.used + __popc(mask & ((1u << lane) - 1u));
  wc.used += cnt;
  return pred ? idx : RPT_NONE;
}
// Binned variant: the warp keeps one chunk per bin (state in shared memory, private to the warp), so the output
// queue becomes a sequence of chunks that each hold rays of ONE bin (direction octant for walk rays, origin cell for
// NEE rays). The consumer's warps then trace rays that take similar routes through the BVH: ncu shows the traversal
// kernels run 30 of 32 lanes on the coherent first bounce but 11-13 on unsorted later bounces.
#define NBINS 8u
#define QCHUNK_BINNED 128u
// per = entries reserved per appending lane (warp-uniform, 32 * per <= CH): the lane owns [idx, idx + per).
template <class Mark>
__device__ __forceinline__ uint32_t chunk_append_binned(uint32_t *counter, WarpChunk *st, bool pred, uint32_t bin, Mark mark, uint32_t CH = QCHUNK_BINNED,
                                                        uint32_t per = 1u) {
  uint32_t active = __ballot_sync(0xFFFFFFFFu, pred);
  uint32_t idx = RPT_NONE;
  if (pred) {
    uint32_t lane = threadIdx.x & 31u;
    uint32_t group = __match_any_sync(active, bin);  // the lanes that append to the same bin
    uint32_t leader = __ffs(group) - 1, cnt = __popc(group) * per;
    uint32_t first = 0;
    if (lane == leader) {
      WarpChunk wc = st[bin];
      if (wc.used + cnt > CH) {
        for (uint32_t e = wc.used; e < CH; ++e) mark(wc.base + e);  // < 32 entries unless the chunk was never used
        wc.base = atomicAdd(counter, CH);
        wc.used = 0;
      }
      first = wc.base + wc.used;
      st[bin] = WarpChunk{wc.base, wc.used + cnt};
    }
    first = __shfl_sync(group, first, leader);
    idx = first + __popc(group & ((1u << lane) - 1u)) * per;
  }
  __syncwarp();
  return idx;
}
template <class Mark>
__device__ __forceinline__ void chunk_pad_binned(const WarpChunk *st, uint32_t nbins, Mark mark, uint32_t CH = QCHUNK_BINNED) {
  uint32_t lane = threadIdx.x & 31u;
  for (uint32_t b = 0; b < nbins; ++b) {
    WarpChunk wc = st[b];
    for (uint32_t e = wc.used + lane; e < CH; e += 32u) mark(wc.base + e);
  }
}

// per-thread statistics -> one atomic per warp at kernel end
__device__ __forceinline__ void flush_count(uint32_t v, uint32_t *dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  if ((threadIdx.x & 31u) == 0 && v) atomicAdd(dst, v);
}
// The same for N counters at once. -DRPT_CTA_FLUSH reduces over the whole CTA first (one atomic per CTA and counter instead
// of one per warp): tried because ~4 700 warps x 3-4 same-line atomics end every launch; measured in one session
// (profiles/r02_flush_variants.txt) it changes nothing on Cornell (20.69 vs 20.67 ms) and costs the all-GGX furnace 4 %
// (the extra __syncthreads at the end of k_nee), so the per-warp form stays the default.
template <int N>
__device__ __forceinline__ void flush_counts_cta(const uint32_t (&v)[N], uint32_t *const (&dst)[N]) {
#ifndef RPT_CTA_FLUSH
#pragma unroll
  for (int k = 0; k < N; ++k) flush_count(v[k], dst[k]);
#else
  __shared__ uint32_t s_cnt[N][32];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nwarps = (blockDim.x + 31u) >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    uint32_t x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
    if (lane == 0) s_cnt[k][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    uint32_t x = 0;
    for (uint32_t w = 0; w < nwarps; ++w) x += s_cnt[threadIdx.x][w];
    if (x) atomicAdd(dst[threadIdx.x], x);
  }
#endif
}

// Dynamic hand-out of a queue's tiles (32 consecutive entries each) to warps. Rays cost 1..30 node visits and vertices
// differ in their NEE work, so with a static split the warps of a launch finish far apart (ncu: achieved occupancy 41 of a
// possible 50 % in the traversal kernels); here a warp claims `grab` tiles with one atomicAdd and keeps claiming until the
// queue is empty, so all warps finish together (Cornell: trace 7.4 -> 5.9 ms, shadow 8.3 -> 7.2, shade 8.2 -> 7.3).
// * grab is sized so that a warp makes about CLAIMS_PER_WARP claims per launch whatever the queue length: the balance
//   granularity stays a few percent of a warp's share, and the claim counter sees a bounded number of atomics (same-address
//   atomics sustain ~0.7 G/s; a fixed grab of 2 tiles cost a streaming scene - one sphere, a ray is one node test - 33 ms
//   of a 27 ms kernel). Scenes whose rays are nearly free (a handful of analytic shapes, no mesh) take at least 4 tiles per
//   claim (DevScene::min_grab): there a claim's latency is long against a tile's work (rtiow2 3.2 -> 2.5 ms per frame), while
//   for real geometry the finer balance wins (Cornell 21.8 vs 22.3 ms, kitchen_sink 7.5 vs 8.7 ms).
// * the NEXT claim is issued as soon as the current one is taken up, so its latency overlaps the tiles being processed.
#define CLAIMS_PER_WARP 32u
struct TileStream {
  uint32_t *ctr;
  uint32_t n_tiles, grab, cur, lim, pending;
  __device__ __forceinline__ uint32_t claim() const { return (threadIdx.x & 31u) == 0 ? atomicAdd(ctr, grab) : 0u; }
  __device__ __forceinline__ void init(uint32_t *counter, uint32_t n_tiles_, uint32_t total_warps, uint32_t min_grab) {
    ctr = counter;
    n_tiles = n_tiles_;
    grab = min(64u, max(min_grab, n_tiles_ / (total_warps * CLAIMS_PER_WARP)));
    cur = lim = 0;
    pending = claim();
  }
  // warp-uniform; returns the next tile or RPT_NONE when the queue is exhausted
  __device__ __forceinline__ uint32_t next() {
    if (cur < lim) return cur++;
    const uint32_t base = __shfl_sync(0xFFFFFFFFu, pending, 0);
    if (base >= n_tiles) return RPT_NONE;  // (pending keeps its value: every later call lands here again)
    cur = base;
    lim = min(base + grab, n_tiles);
    pending = claim();
    return cur++;
  }
};

#define TRACE_THREADS 128
#ifndef TRACE_DYNAMIC
#define TRACE_DYNAMIC 1
#endif
#ifndef TRACE_MIN_BLOCKS
#define TRACE_MIN_BLOCKS 8  // resident CTAs per SM the traversal kernels are compiled for (register cap = 65536 / (128 * this))
#endif
#ifndef TRACE_MIN_BLOCKS_FLAT
#define TRACE_MIN_BLOCKS_FLAT TRACE_MIN_BLOCKS  // the one-level walk (TRAV_BVH_FLAT) needs no spills at 64 registers; tighter caps: profiles/r02_flat_walk.txt
#endif
#ifndef TRACE_MIN_BLOCKS_WIDE
#define TRACE_MIN_BLOCKS_WIDE 6  // the four-wide walk holds seven 128-bit node rows in flight: 80 registers instead of 64
#endif

// per-thread BVH work counters -> one 64-bit atomic per warp per counter at kernel end
__device__ __forceinline__ void flush_work(const TraceWork &w, unsigned long long *work) {
  uint32_t a = w.nodes, b = w.tris, c = w.insts;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
    b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
    c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  }
  if ((threadIdx.x & 31u) == 0) {
    atomicAdd(work + 0, (unsigned long long)a);
    atomicAdd(work + 1, (unsigned long long)b);
    atomicAdd(work + 2, (unsigned long long)c);
  }
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// The camera vertex of sample `slot` (pt.rs:397-446): film jitter, wavelength, camera ray.
__device__ __forceinline__ PathRec camera_record(const RenderCtx &R, uint32_t slot) {
    uint32_t pixel = slot_to_pixel(R, slot % R.wh), sample = R.sample_base + slot / R.wh;
    uint32_t px = pixel % R.width, py = pixel / R.width;
    RptRand4 s0 = rpt_philox(R.seed, pixel, sample, 0);
    RptRand4 s1 = rpt_philox(R.seed, pixel, sample, 1);
    // box filter (tiled.rs:372-375), wavelength (pt.rs:406), film clamp (pt.rs:411-414)
    float cu = ((float)px + s0.x) / (float)R.width, cv = ((float)py + s0.y) / (float)R.height;
    float lambda = R.lambda_lo + s0.z * (R.lambda_hi - R.lambda_lo);
    float fu = clampf(cu, 0.0f, 1.0f - RPT_EPS), fv = clampf(cv, 0.0f, 1.0f - RPT_EPS);
    float3 origin, dir;
    float3 cu3 = f3(R.cam.u[0], R.cam.u[1], R.cam.u[2]), cv3 = f3(R.cam.v[0], R.cam.v[1], R.cam.v[2]);
    if (R.cam.kind == RPT_CAMERA_PANORAMA) {
      // PanoramaCamera::get_ray (camera/panorama_camera.rs:68-91): pinhole, direction from azimuth / elevation; not renormalised
      float ax = R.cam.angle_span[0] * (fu - 0.5f), ay = R.cam.angle_span[1] * (0.5f - fv);
      float sx, cx, sy, cy;
      sincosf(ax, &sx, &cx);
      sincosf(ay, &sy, &cy);
      float3 vec = f3(sx * cy, sy, cx * cy);
      origin = f3(R.cam.origin[0], R.cam.origin[1], R.cam.origin[2]);
      dir = cu3 * vec.x + cv3 * vec.y + f3(R.cam.w[0], R.cam.w[1], R.cam.w[2]) * vec.z;
    } else {
      // ProjectiveCamera::get_ray (camera/projective_camera.rs:101-120)
      float3 vec = random_in_unit_disk(s1.x, s1.y);
      float3 rd = R.cam.aperture_diameter * vec;
      origin = f3(R.cam.origin[0], R.cam.origin[1], R.cam.origin[2]) + (cu3 * rd.x + cv3 * rd.y);
      float3 pop = f3(R.cam.lower_left[0], R.cam.lower_left[1], R.cam.lower_left[2]) +
                   fu * f3(R.cam.horizontal[0], R.cam.horizontal[1], R.cam.horizontal[2]) +
                   fv * f3(R.cam.vertical[0], R.cam.vertical[1], R.cam.vertical[2]);
      dir = normalized(pop - origin);
    }
    PathRec r;
    r.r0 = make_float4(origin.x, origin.y, origin.z, 1.0f);
    r.r1 = make_float4(dir.x, dir.y, dir.z, 100.0f);
    r.r2 = make_float4(dir.x, dir.y, dir.z, lambda);
    r.r3 = make_float4(__uint_as_float(slot), 0.0f, 0.0f, 0.0f);
    return r;
}

__global__ void __launch_bounds__(256) k_raygen(DevScene S, RenderCtx R, PathRec *__restrict__ out, uint32_t *__restrict__ counts) {
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < R.n_slots; slot += stride) out[slot] = camera_record(R, slot);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[Q_PATHS] = R.n_slots;
    counts[N_PATHS] = R.n_slots;
  }
}

__device__ __forceinline__ uint32_t material_class(const DevScene &S, uint32_t material) {
  return S.materials[RPT_MAT_INDEX(material)].type == RPT_MATERIAL_GGX ? Q_GGX : Q_DIFFUSE;
}

// Closest-hit traversal of the path queue; appends each path to the list of its vertex's class.
// RAYGEN: the launch of bounce 0 generates its camera vertices itself (and writes them out for the shade kernel) instead
// of reading what a separate ray-generation kernel wrote: one 64-byte queue write + read per sample less.
enum : int { TRAV_BVH = 0, TRAV_BVH_TMA = 1, TRAV_SMALL = 2, TRAV_BVH_REFILL = 3, TRAV_BVH4 = 4, TRAV_BVH_FLAT = 5 };  // how the traversal kernels find hits (chosen per scene)

// TRAV_BVH_FLAT: TRAV_BVH for scenes without a transformed mesh instance (TravT<., TWO_LEVEL = false>); chosen automatically,
// RPT_FLAT=0 keeps the two-level walk.
// TRAV_BVH4 (RPT_BVH4=1): the same trees collapsed to four children per node (rpt::collapse_bvh4, TravT<true>): half the
// dependent node fetches per ray, one 128-byte line per node, no per-box min / max.

// TRAV_BVH_REFILL (opt-in, RPT_REFILL=1): lanes whose ray has finished are re-armed with the next ray of the queue while the
// other lanes of the warp are still walking ("persistent threads with dynamic fetch", Aila & Laine 2009), once at least
// REFILL_MIN lanes are idle. Motivation: on the 10 M-triangle instanced scene every part of the traversal kernels - node
// loop, leaf code, instance entry - runs at 6.5-7.7 of 32 lanes (ncu source view): rays take 5 to 100+ node visits, a warp
// lives as long as its longest ray, and the lanes of the short rays idle. Outcome (profiles/r02_refill_vs_tile.md): lanes per
// instruction rise to 9.8-12.8, frame time does not move (213 vs 214 ms); see rpt_scene_create for why it is not the default.
#ifndef REFILL_MIN
#define REFILL_MIN 8u
#endif

// Copies the small-scene triangle table into the CTA's dynamic shared memory (TRAV_SMALL only).
__device__ __forceinline__ void stage_small_tris(const DevScene &S, float4 *s_tris) {
  for (uint32_t i = threadIdx.x; i < 9u * S.small_ntri; i += blockDim.x) s_tris[i] = __ldg(S.small_tris + i);
  __syncthreads();
}

template <int MODE, bool RAYGEN, bool STATS>
__global__ void __launch_bounds__(TRACE_THREADS, MODE == 4 ? TRACE_MIN_BLOCKS_WIDE : (MODE == 5 ? TRACE_MIN_BLOCKS_FLAT : TRACE_MIN_BLOCKS)) k_trace(DevScene S, const PathRec *__restrict__ paths, HitRec *__restrict__ hits,
                                                         uint32_t *__restrict__ q_miss, uint32_t *__restrict__ q_diffuse,
                                                         uint32_t *__restrict__ q_ggx, uint32_t *__restrict__ counts,
                                                         unsigned long long *__restrict__ work, float *__restrict__ acc,
                                                         RenderCtx R, PathRec *__restrict__ paths_out) {
  constexpr bool TMA = MODE == TRAV_BVH_TMA;
  static_assert(!(TMA && RAYGEN), "the fused ray generation has no input queue to stage");
  extern __shared__ int s_stack[];  // BVH: [stack entry][thread], depth chosen per scene at rpt_scene_create; SMALL: the triangle table
  if (MODE == TRAV_SMALL) stage_small_tris(S, reinterpret_cast<float4 *>(s_stack));
  // Queue read, two selectable forms (rpt_scene_create picks one; RPT_TMA_TILES=1 selects the TMA form):
  //  * TMA-staged tiles: each warp owns two 2 KB buffers (32 path records each) and two mbarriers. Lane 0 arms the
  //    barrier with the tile's byte count and issues one cp.async.bulk (UBLKCP) for the NEXT tile; the warp traces
  //    the current tile out of shared memory.
  //  * plain coalesced 128-bit loads of the warp's tile.
  // Measured on the bench workload the two are within 5 % (TMA slightly slower): the traversal is issue-bound and
  // the queue read is ~3 % of its stall samples, so there is no latency for the staging to hide.
  constexpr int kWarps = TRACE_THREADS / 32;
  __shared__ __align__(128) float4 s_tile[TMA ? kWarps : 1][2][TMA ? 32 * 4 : 1];
  __shared__ __align__(8) uint64_t s_bar[kWarps][2];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  if (TMA) {
    if (threadIdx.x < kWarps * 2) mbar_init(&s_bar[threadIdx.x >> 1][threadIdx.x & 1], 1);
    mbar_fence_init();
    __syncthreads();
  }
  TraceWork tw{0, 0, 0};
  const uint32_t n = RAYGEN ? R.n_slots : counts[Q_PATHS];
  const uint32_t qc = chunk_size_for(n);
  WarpChunk wc_miss = chunk_init(qc), wc_diffuse = chunk_init(qc), wc_ggx = chunk_init(qc);
  uint32_t n_miss = 0, n_diffuse = 0, n_ggx = 0;
  if (RAYGEN && blockIdx.x == 0 && threadIdx.x == 0) {
    counts[Q_PATHS] = n;
    counts[N_PATHS] = n;
  }
  const uint32_t n_tiles = (n + 31u) >> 5;
  const uint32_t total_warps = gridDim.x * kWarps;
  auto issue = [&](uint32_t stage, uint32_t tile) {
    uint32_t bytes = min(32u, n - tile * 32u) * (uint32_t)sizeof(PathRec);
    mbar_expect_tx(&s_bar[warp][stage], bytes);
    tma_load_1d(&s_tile[TMA ? warp : 0][stage][0], paths + (size_t)tile * 32u, bytes, &s_bar[warp][stage]);
  };
  uint32_t stage = 0, phase = 0;  // phase bit k = parity to wait for on stage k
  uint32_t tile0 = blockIdx.x * kWarps + warp;
  if (TMA && tile0 < n_tiles && lane == 0) issue(0, tile0);
  auto body = [&](uint32_t tile) {
    const uint32_t i = tile * 32u + lane;
    bool active = i < n;
    uint32_t cls = RPT_NONE;
    PathRec r;
    if (TMA) {
      uint32_t next = tile + total_warps;
      if (next < n_tiles && lane == 0) issue(stage ^ 1u, next);
      mbar_wait(&s_bar[warp][stage], (phase >> stage) & 1u);
      phase ^= 1u << stage;
      if (active) {
        const float4 *rp = &s_tile[TMA ? warp : 0][stage][lane * 4];
        r.r0 = rp[0];
        r.r1 = rp[1];
        r.r2 = rp[2];
        r.r3 = rp[3];
      }
      __syncwarp();  // every lane holds its record in registers: the buffer may be refilled
      stage ^= 1u;
    } else if (RAYGEN) {
      if (active) {
        r = camera_record(R, i);
        paths_out[i] = r;
      }
    } else if (active) {
      const float4 *rp = reinterpret_cast<const float4 *>(paths + i);
      r.r0 = __ldg(rp);
      r.r1 = __ldg(rp + 1);
      r.r2 = __ldg(rp + 2);
      r.r3 = __ldg(rp + 3);
    }
    if (active) active = __float_as_uint(r.r3.x) != RPT_NONE;  // padding of an abandoned chunk tail
    TraceHit th;
    bool hit = false;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);
    if (active) {
      o = rec_origin(r);
      d = f3(r.r2);
    }
    if (MODE == TRAV_SMALL) {
      SmallTrav sv;
      sv.init(RPT_INF);
      sv.template run<false, STATS>(S, reinterpret_cast<const float4 *>(s_stack), o, d, RPT_INF, active, tw);
      hit = sv.found;
      th = sv.out;
    } else if (active) {
      hit = trace_ray<false, STATS, MODE == TRAV_BVH4, MODE != TRAV_BVH_FLAT>(S, o, d, RPT_INF, s_stack + threadIdx.x, TRACE_THREADS, th, tw);
    }
    if (active) {
      HitRec h;
      h.t = th.t;
      h.inst = th.inst;
      h.prim = th.prim;
      h.mat = RPT_NONE;
      if (!hit) {
        if (S.env_kind == RPT_ENV_CONSTANT) {
          // The environment vertex of a direction-independent environment is finished here, from the path record that is
          // still in registers (same arithmetic as k_shade_miss): no class-list entry, no 64-byte re-gather.
          float lambda = r.r2.w, beta = r.r0.w, pdf_fwd = r.r1.w;
          float emission = curve_eval(S, S.env_curve, lambda) * S.env_strength;  // env_emission, Constant
          float cos_i = fabsf(dot(f3(r.r1), d));
          float nee_psa_pdf = (1.0f / (4.0f * RPT_PI)) / fabsf(cos_i);  // env_pdf_for, Constant
          float bsdf_psa_pdf = pdf_fwd / fabsf(cos_i);
          float c = power_heuristic(bsdf_psa_pdf, nee_psa_pdf) * beta * emission;
          if (c != 0.0f) atomicAdd(acc + __float_as_uint(r.r3.x), c);
          n_miss++;
        } else {
          cls = Q_MISS;
        }
      } else {
        const DevInstance &I = S.instances[th.inst];
        uint32_t mat = I.material;
        if (mat == RPT_NONE) {
          mat = RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
          if ((I.flags & DI_KIND_MASK) == RPT_AGG_MESH) mat = __float_as_uint(__ldg(S.tri_verts + 3 * (size_t)(I.tri_base + th.prim)).w);
        }
        cls = material_class(S, mat);
        h.mat = mat;
      }
      hits[i] = h;
    }
    uint32_t k;
    k = chunk_append(counts + Q_MISS, wc_miss, cls == Q_MISS, [&](uint32_t e) { q_miss[e] = RPT_NONE; }, qc);
    if (cls == Q_MISS) q_miss[k] = i;
    k = chunk_append(counts + Q_DIFFUSE, wc_diffuse, cls == Q_DIFFUSE, [&](uint32_t e) { q_diffuse[e] = RPT_NONE; }, qc);
    if (cls == Q_DIFFUSE) q_diffuse[k] = i;
    k = chunk_append(counts + Q_GGX, wc_ggx, cls == Q_GGX, [&](uint32_t e) { q_ggx[e] = RPT_NONE; }, qc);
    if (cls == Q_GGX) q_ggx[k] = i;
    n_miss += cls == Q_MISS;
    n_diffuse += cls == Q_DIFFUSE;
    n_ggx += cls == Q_GGX;
  };
  if (MODE == TRAV_BVH_REFILL) {
    const PathRec *src = RAYGEN ? paths_out : paths;
    TileStream ts;
    ts.init(counts + F_TRACE, n_tiles, total_warps, S.min_grab);
    uint32_t pool_cur = 0, pool_end = 0, my_i = 0;  // pool: the rays of the tile the warp is handing out (warp-uniform)
    bool has = false, exhausted = false;
    Trav t;
    const uint32_t lt_mask = (1u << lane) - 1u;
    while (true) {
      const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !has);
      if (!exhausted && (__popc(idle) >= REFILL_MIN || idle == 0xFFFFFFFFu)) {
        const uint32_t need = __popc(idle), rank = __popc(idle & lt_mask);
        uint32_t given = 0;
        while (given < need) {
          if (pool_cur == pool_end) {
            const uint32_t tile = ts.next();
            if (tile == RPT_NONE) {
              exhausted = true;
              break;
            }
            pool_cur = tile * 32u;
            pool_end = min(pool_cur + 32u, n);
          }
          const uint32_t take = min(need - given, pool_end - pool_cur);
          if (!has && rank >= given && rank < given + take) {
            my_i = pool_cur + (rank - given);
            PathRec r;
            if (RAYGEN) {
              r = camera_record(R, my_i);
              paths_out[my_i] = r;
            } else {
              const float4 *rp = reinterpret_cast<const float4 *>(paths + my_i);
              r.r0 = __ldg(rp);
              r.r1 = __ldg(rp + 1);
              r.r2 = __ldg(rp + 2);
              r.r3 = __ldg(rp + 3);
            }
            if (__float_as_uint(r.r3.x) != RPT_NONE) {  // (padding of an abandoned chunk tail: the lane stays idle until the next hand-out)
              t.init(S, rec_origin(r), f3(r.r2), RPT_INF);
              has = true;
            }
          }
          pool_cur += take;
          given += take;
        }
      }
      if (!__any_sync(0xFFFFFFFFu, has)) {
        if (exhausted) break;
        continue;  // the hand-out met only padding: take the next rays
      }
      bool finished = false;
      if (has) finished = t.template step<false, STATS>(S, s_stack + threadIdx.x, TRACE_THREADS, tw);
      if (__any_sync(0xFFFFFFFFu, finished)) {
        uint32_t cls = RPT_NONE;
        if (finished) {
          has = false;
          HitRec h;
          h.t = t.out.t;
          h.inst = t.out.inst;
          h.prim = t.out.prim;
          h.mat = RPT_NONE;
          if (!t.found) {
            if (S.env_kind == RPT_ENV_CONSTANT) {  // as in the tile form below, from the re-read path record
              const float4 *rp = reinterpret_cast<const float4 *>(src + my_i);
              float4 r0 = rp[0], r1 = rp[1], r2 = rp[2], r3 = rp[3];
              float lambda = r2.w, beta = r0.w, pdf_fwd = r1.w;
              float emission = curve_eval(S, S.env_curve, lambda) * S.env_strength;
              float cos_i = fabsf(dot(f3(r1), t.d));
              float nee_psa_pdf = (1.0f / (4.0f * RPT_PI)) / fabsf(cos_i);
              float bsdf_psa_pdf = pdf_fwd / fabsf(cos_i);
              float c = power_heuristic(bsdf_psa_pdf, nee_psa_pdf) * beta * emission;
              if (c != 0.0f) atomicAdd(acc + __float_as_uint(r3.x), c);
              n_miss++;
            } else {
              cls = Q_MISS;
            }
          } else {
            const DevInstance &I = S.instances[t.out.inst];
            uint32_t mat = I.material;
            if (mat == RPT_NONE) {
              mat = RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
              if ((I.flags & DI_KIND_MASK) == RPT_AGG_MESH) mat = __float_as_uint(__ldg(S.tri_verts + 3 * (size_t)(I.tri_base + t.out.prim)).w);
            }
            cls = material_class(S, mat);
            h.mat = mat;
          }
          hits[my_i] = h;
        }
        uint32_t k;
        k = chunk_append(counts + Q_MISS, wc_miss, cls == Q_MISS, [&](uint32_t e) { q_miss[e] = RPT_NONE; }, qc);
        if (cls == Q_MISS) q_miss[k] = my_i;
        k = chunk_append(counts + Q_DIFFUSE, wc_diffuse, cls == Q_DIFFUSE, [&](uint32_t e) { q_diffuse[e] = RPT_NONE; }, qc);
        if (cls == Q_DIFFUSE) q_diffuse[k] = my_i;
        k = chunk_append(counts + Q_GGX, wc_ggx, cls == Q_GGX, [&](uint32_t e) { q_ggx[e] = RPT_NONE; }, qc);
        if (cls == Q_GGX) q_ggx[k] = my_i;
        n_miss += cls == Q_MISS;
        n_diffuse += cls == Q_DIFFUSE;
        n_ggx += cls == Q_GGX;
      }
    }
  } else if (TMA || (!TRACE_DYNAMIC && MODE != TRAV_SMALL)) {
    for (uint32_t tile = tile0; tile < n_tiles; tile += total_warps) body(tile);
  } else {
    TileStream ts;
    ts.init(counts + F_TRACE, n_tiles, total_warps, S.min_grab);
    for (uint32_t tile = ts.next(); tile != RPT_NONE; tile = ts.next()) body(tile);
  }
  if (wc_miss.used < qc) chunk_pad(wc_miss, [&](uint32_t e) { q_miss[e] = RPT_NONE; }, qc);
  if (wc_diffuse.used < qc) chunk_pad(wc_diffuse, [&](uint32_t e) { q_diffuse[e] = RPT_NONE; }, qc);
  if (wc_ggx.used < qc) chunk_pad(wc_ggx, [&](uint32_t e) { q_ggx[e] = RPT_NONE; }, qc);
  {
    const uint32_t v[3] = {n_miss, n_diffuse, n_ggx};
    uint32_t *const d[3] = {counts + N_MISS, counts + N_DIFFUSE, counts + N_GGX};
    flush_counts_cta<3>(v, d);
  }
  if (STATS) flush_work(tw, work);
}

// Environment vertex (integrator/utils.rs:344-373 + pt.rs:487-511).
__global__ void __launch_bounds__(256) k_shade_miss(DevScene S, const PathRec *__restrict__ paths, const uint32_t *__restrict__ queue,
                                                    const uint32_t *__restrict__ counts, float *__restrict__ acc) {
  const uint32_t n = counts[Q_MISS];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t idx = queue[i];
    if (idx == RPT_NONE) continue;  // chunk padding
    const float4 *rp = reinterpret_cast<const float4 *>(paths + idx);
    float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2), r3 = __ldg(rp + 3);
    float3 wo = f3(r2);
    float lambda = r2.w, beta = r0.w, pdf_fwd = r1.w;
    float u, v;
    direction_to_uv(wo, u, v);
    float emission = env_emission(S, u, v, lambda);
    float cos_i = fabsf(dot(f3(r1), wo));
    float nee_psa_pdf = env_pdf_for(S, u, v) / fabsf(cos_i);
    float bsdf_psa_pdf = pdf_fwd / fabsf(cos_i);
    float weight = power_heuristic(bsdf_psa_pdf, nee_psa_pdf);
    float c = weight * beta * emission;
    if (c != 0.0f) atomicAdd(acc + __float_as_uint(r3.x), c);
  }
}

// One walk vertex of a material class: light-hit MIS (pt.rs:512-561), NEE generation
// (pt.rs:562-604,333-393,146-219,224-331), BSDF sampling + russian roulette (integrator/utils.rs:214-329).
#define SHADE_THREADS 128
#ifndef BIN_MIN_ITEMS
#define BIN_MIN_ITEMS (4u << 20)
#endif
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 7  // 72 registers: measured 6 vs 7 CTAs/SM: Cornell shade 7.40 -> 7.33 ms, 4K HDR scene 351 -> 333 ms, gem +1 %
#endif
#ifndef SHADE_DYNAMIC
#define SHADE_DYNAMIC 1
#endif
template <uint32_t CLASS>
__global__ void __launch_bounds__(SHADE_THREADS, SHADE_MIN_BLOCKS) k_shade_surface(DevScene S, RenderCtx R, uint32_t bounce, const PathRec *__restrict__ paths,
                                                       const HitRec *__restrict__ hits, const uint32_t *__restrict__ queue,
                                                       uint32_t *__restrict__ counts, uint32_t *__restrict__ next_counts,
                                                       PathRec *__restrict__ out, float4 *__restrict__ sh_a, float4 *__restrict__ sh_b,
                                                       uint32_t *__restrict__ sh_c, float *__restrict__ acc) {
  // Software-pipelined gather: while item i is shaded, the path + hit record of item i + stride (80 B, reached
  // through the class list) is already in flight to this thread's private shared-memory slot via cp.async, so the
  // DRAM latency of the gather overlaps the shading arithmetic instead of stalling it (ncu v1: long-scoreboard
  // 10 warps per issue). [stage][float4 k][thread]: consecutive threads hit consecutive 16-byte words.
  __shared__ float4 s_stage[2][5][SHADE_THREADS];
  const uint32_t n = counts[CLASS];
  const uint32_t n_round = (n + 31u) & ~31u;
  const uint32_t L = R.light_samples;
  const uint32_t max_bounces = R.only_direct ? 1u : R.max_bounces;
  const uint32_t tid = threadIdx.x;
  auto issue = [&](int stage, uint32_t idx) {
    if (idx != RPT_NONE) {
      const float4 *rp = reinterpret_cast<const float4 *>(paths + idx);
#pragma unroll
      for (int k = 0; k < 4; ++k) cp_async16(&s_stage[stage][k][tid], rp + k);
      cp_async16(&s_stage[stage][4][tid], hits + idx);
    }
    cp_async_commit();
  };
  // shadow rays are binned by origin cell: per-warp chunk states [warp][bin] in shared memory; the next-path queue
  // keeps a single register-resident chunk (binning walk rays by direction octant was measured: no gain)
  __shared__ WarpChunk s_chunks[SHADE_THREADS / 32][NBINS];
  WarpChunk *st_shadow = s_chunks[threadIdx.x >> 5];
  const uint32_t qc = chunk_size_for(n), bc = n >= BIN_MIN_ITEMS ? QCHUNK_BINNED : QCHUNK_BINNED_SMALL;
  WarpChunk wc_next = chunk_init(qc);
  if ((threadIdx.x & 31u) < NBINS) st_shadow[threadIdx.x & 31u] = WarpChunk{0u, bc};
  __syncwarp();
  const uint32_t nbins = n >= BIN_MIN_ITEMS ? NBINS : 1u;  // small queues: binning would only scatter a few rays over many chunks
  uint32_t n_next = 0, n_shadow = 0, n_sh_ref = 0, n_nan = 0;
  auto mark_next = [&](uint32_t e) { out[e].r3 = make_float4(__uint_as_float(RPT_NONE), 0.0f, 0.0f, 0.0f); };
  auto mark_shadow = [&](uint32_t e) { sh_c[e] = RPT_NONE; };
  // The warp's stream of tiles (32 consecutive queue entries each): claimed from a device counter through TileStream
  // (or, SHADE_DYNAMIC=0, the static round-robin split), two tiles ahead of the one being shaded so that the cp.async
  // prefetch pipeline below never drains at a claim boundary.
  const uint32_t lane = tid & 31u;
  const uint32_t n_tiles = n_round >> 5;
#if SHADE_DYNAMIC
  TileStream ts;
  ts.init(counts + (CLASS == Q_DIFFUSE ? F_SHADE_DIFFUSE : F_SHADE_GGX), n_tiles, gridDim.x * (SHADE_THREADS / 32), S.min_grab);
  auto next_tile = [&]() -> uint32_t { return ts.next(); };
#else
  uint32_t g_cur = blockIdx.x * (SHADE_THREADS / 32) + (tid >> 5);
  auto next_tile = [&]() -> uint32_t {
    uint32_t t = g_cur;
    g_cur += gridDim.x * (SHADE_THREADS / 32);
    return t < n_tiles ? t : RPT_NONE;
  };
#endif
  auto tile_idx = [&](uint32_t t) -> uint32_t {
    uint32_t i = t * 32u + lane;
    return (t != RPT_NONE && i < n) ? __ldg(queue + i) : RPT_NONE;
  };
  uint32_t t_cur = next_tile(), t_next = next_tile();
  uint32_t idx_cur = tile_idx(t_cur), idx_next = tile_idx(t_next);
  issue(0, idx_cur);
  int stage = 0;
  while (t_cur != RPT_NONE) {
    issue(stage ^ 1, idx_next);  // prefetch the next item while this one is shaded
    uint32_t t_nn = next_tile();
    uint32_t idx_nn = tile_idx(t_nn);
    cp_async_wait<1>();  // everything but the group just committed has landed
    bool active = idx_cur != RPT_NONE;  // beyond the queue, or chunk padding
    bool continues = false, do_nee = false;
    PathRec nr;
    // state shared between the vertex evaluation and the warp-uniform NEE loop below
    SurfaceHit sh;
    Frame frame;
    float3 wi_nee = f3(0, 0, 1);
    float beta = 0.0f, lambda = 0.0f, albedo = 0.0f;
    uint32_t slot = 0, pixel = 0, sample = 0;
    GgxParams gp;
    gp.alpha = 1.0f;
    gp.eta_inner = gp.eta_outer = 1.0f;
    gp.kappa = 0.0f;
    gp.metallic = false;
    if (active) {
      PathRec r;
      r.r0 = s_stage[stage][0][tid];
      r.r1 = s_stage[stage][1][tid];
      r.r2 = s_stage[stage][2][tid];
      r.r3 = s_stage[stage][3][tid];
      float4 hraw = s_stage[stage][4][tid];
      TraceHit th;
      th.t = hraw.x;
      th.inst = __float_as_uint(hraw.y);
      th.prim = __float_as_uint(hraw.z);
      float3 prev_p = f3(r.r0), prev_n = f3(r.r1), d = f3(r.r2);
      float prev_pdf = r.r1.w;
      beta = r.r0.w;
      lambda = r.r2.w;
      slot = __float_as_uint(r.r3.x);
      float3 o = rec_origin(r);
      reconstruct_hit(S, o, d, th, sh);
      frame = frame_from_normal(sh.n);
      float3 wi = normalized(to_local(frame, -d));  // integrator/utils.rs:175-176
      const RptMaterial m = S.materials[RPT_MAT_INDEX(sh.material)];
      pixel = slot_to_pixel(R, slot % R.wh);
      sample = R.sample_base + slot / R.wh;
      RptRand4 s = rpt_philox(R.seed, pixel, sample, rpt_block_bsdf(bounce, L));

      // ---- generate_and_evaluate
      float3 wo;
      float f, pdf;
      if (CLASS == Q_GGX) {
        gp = ggx_params(S, m, lambda);
        Bsdf b = ggx_generate_and_evaluate(gp, s.x, s.y, wi, wo);
        f = b.f;
        pdf = b.pdf;
      } else {
        albedo = diffuse_albedo(S, m, lambda, sh.u, sh.v);
        wo = random_cosine_direction(s.x, s.y) * signumf(wi.z);
        f = albedo / RPT_PI;
        pdf = fabsf(wo.z) / RPT_PI;
      }
#ifdef RPT_DEBUG
      if (slot == RPT_DEBUG_SLOT)
        printf("[shade b%u cls%u] p=(%.7g %.7g %.7g) n=(%.7g %.7g %.7g) mat=%u beta=%.7g prev_pdf=%.7g f=%.7g pdf=%.7g wo=(%.7g %.7g %.7g) wi=(%.7g %.7g %.7g) s=(%.7g %.7g %.7g) albedo=%.7g inst=%u prim=%u t=%.7g\n",
               bounce, CLASS, sh.p.x, sh.p.y, sh.p.z, sh.n.x, sh.n.y, sh.n.z, sh.material, beta, prev_pdf, f, pdf, wo.x, wo.y, wo.z, wi.x, wi.y, wi.z, s.x, s.y, s.z, albedo, th.inst, th.prim, th.t);
#endif
      if (pdf != pdf) {
        // pdf NaN: the walk breaks BEFORE pushing the vertex (integrator/utils.rs:261-263)
        n_nan++;
      } else {
        // ---- the vertex exists: its contribution (second loop of pt.rs:481-613)
        if (RPT_MAT_IS_LIGHT(sh.material)) {
          float emission = material_emission(S, m, lambda, wi);
          if (emission > 0.0f) {
            float c = 0.0f;
            if (L == 0 || bounce == 0) {
              c = beta * emission;
            } else if (!R.only_direct) {
              float3 nee_direction = normalized(sh.p - prev_p);
              float hyp = instance_psa_pdf(S.instances[th.inst], dot(prev_n, nee_direction), dot(sh.n, nee_direction), prev_p, sh.p);
              c = power_heuristic(prev_pdf, hyp) * beta * emission;
            }
#ifdef RPT_DEBUG
            if (slot == RPT_DEBUG_SLOT) printf("[emit b%u] emission=%.7g c=%.7g\n", bounce, emission, c);
#endif
            if (c != 0.0f) atomicAdd(acc + slot, c);
          }
        } else if (L > 0) {
          do_nee = true;
          wi_nee = to_local(frame, normalized(prev_p - sh.p));  // pt.rs:565-569
        }
        // ---- continue the walk (integrator/utils.rs:266-329)
        float cos_o = fabsf(wo.z);
        float rr = bounce >= R.min_bounces ? fminf(f / pdf, 1.0f) : 1.0f;
        float pdf_forward = pdf * (rr / cos_o);
        float nbeta = beta * (f / pdf_forward);
        if (pdf_forward == 0.0f) nbeta = 0.0f;
        if (nbeta != 0.0f && !(s.z > rr) && bounce + 1 < max_bounces) {
          float3 nd = normalized(to_world(frame, wo));
          nr.r0 = make_float4(sh.p.x, sh.p.y, sh.p.z, nbeta);
          nr.r1 = make_float4(sh.n.x, sh.n.y, sh.n.z, pdf_forward);
          nr.r2 = make_float4(nd.x, nd.y, nd.z, lambda);
          nr.r3 = make_float4(__uint_as_float(slot), signumf(wo.z), 0.0f, 0.0f);
          continues = true;
        }
      }
    }
    // ---- next path queue: one atomic per warp
    uint32_t k = chunk_append(next_counts + Q_PATHS, wc_next, continues, mark_next, qc);
    if (continues) out[k] = nr;
    n_next += continues;

    // ---- NEE: warp-uniform loop over the light samples; every iteration compacts the lanes that
    // produced a shadow ray into the shadow queue with one atomic (estimate_direct_illumination_with_loop).
    if (__any_sync(0xFFFFFFFFu, do_nee)) {
      float inv_l = 1.0f / (float)L;
      for (uint32_t ls = 0; ls < L; ++ls) {
        bool has = false;
        float4 a = make_float4(0, 0, 0, 0), b4 = make_float4(0, 0, 0, 0);
        uint32_t c = 0;
        if (do_nee) {
          RptRand4 sn = rpt_philox(R.seed, pixel, sample, rpt_block_nee(bounce, L, ls));
          float pick;
          bool sample_world = choose(sn.x, S.p_env, pick);  // pt.rs:350-353
          float3 dir = f3(0, 0, 1);
          float light_pdf = 0.0f, emission = 1.0f;
          bool valid = true;
          if (sample_world) {
            float eu, ev;
            env_sample_uv(S, sn.y, sn.z, eu, ev, light_pdf);
            dir = uv_to_direction(eu, ev);
            emission = env_emission(S, eu, ev, lambda);
          } else {
            valid = S.num_lights > 0;
            if (valid) {
              uint32_t li = (uint32_t)clampf((float)S.num_lights * pick, 0.0f, (float)S.num_lights - 1.0f);  // world/mod.rs:109
              instance_sample(S.instances[S.lights[li]], sn.y, sn.z, sh.p, dir, light_pdf);
              light_pdf = light_pdf * (1.0f / (float)S.num_lights);
              valid = light_pdf != 0.0f;  // pt.rs:151-153
            }
          }
          float3 local_wo = to_local(frame, dir);
          if (sample_world && local_wo.z <= 0.0f) valid = false;  // pt.rs:245-247
          if (valid) {
            Bsdf bs;
            if (CLASS == Q_GGX) {
              bs = ggx_bsdf(gp, wi_nee, local_wo);
            } else {  // lambertian.rs:16-32 / diffuse_light.rs:29-45
              bool same = local_wo.z * wi_nee.z > 0.0f;
              bs.f = same ? albedo / RPT_PI : 0.0f;
              bs.pdf = same ? fabsf(local_wo.z) / RPT_PI : 0.0f;
            }
            n_sh_ref++;  // the reference traces (and counts) this ray whatever its weight
            float weight = R.only_direct ? 1.0f : power_heuristic_generic(light_pdf, bs.pdf);
            float pre;
            float3 so;
            if (sample_world) {
              pre = beta * weight * bs.f * emission * fabsf(local_wo.z) * (1.0f / light_pdf) * inv_l;  // pt.rs:313-318
              so = sh.p + (sh.n * RPT_NORMAL_OFFSET) * signumf(dir.z);  // WORLD z (quirk Q12, pt.rs:256)
              c = slot | 0x80000000u;
            } else {
              pre = bs.f * beta * fabsf(local_wo.z) * weight / light_pdf * inv_l;  // pt.rs:196-202 minus the light-side terms
              so = sh.p + (sh.n * RPT_NORMAL_OFFSET) * signumf(local_wo.z);  // pt.rs:171-174
              c = slot;
            }
#ifdef RPT_DEBUG
            if (slot == RPT_DEBUG_SLOT)
              printf("[nee b%u k%u] world=%d dir=(%.7g %.7g %.7g) light_pdf=%.7g bsdf f=%.7g pdf=%.7g weight=%.7g pre=%.7g local_wo.z=%.7g\n", bounce, ls,
                     (int)sample_world, dir.x, dir.y, dir.z, light_pdf, bs.f, bs.pdf, weight, pre, local_wo.z);
#endif
            if (pre != 0.0f) {  // a zero pre-factor cannot contribute: skip the visibility query
              has = true;
              a = make_float4(so.x, so.y, so.z, pre);
              b4 = make_float4(dir.x, dir.y, dir.z, lambda);
            }
          }
        }
        uint32_t bin_sh = nbins > 1 ? ((a.x < S.world_center.x) | ((a.y < S.world_center.y) << 1) | ((a.z < S.world_center.z) << 2)) : 0u;  // origin cell
        uint32_t q = chunk_append_binned(counts + Q_SHADOW, st_shadow, has, bin_sh, mark_shadow, bc);
        if (has) {
          sh_a[q] = a;
          sh_b[q] = b4;
          sh_c[q] = c;
        }
        n_shadow += has;
      }
    }
    stage ^= 1;
    idx_cur = idx_next;
    idx_next = idx_nn;
    t_cur = t_next;
    t_next = t_nn;
  }
  cp_async_wait<0>();
  if (wc_next.used < qc) chunk_pad(wc_next, mark_next, qc);
  chunk_pad_binned(st_shadow, NBINS, mark_shadow, bc);
  {
    const uint32_t v[4] = {n_next, n_shadow, n_sh_ref, n_nan};  // Q_SHADOW_REF: reference-definition shadow-ray counter (pt.rs:176,252)
    uint32_t *const d[4] = {next_counts + N_PATHS, counts + N_SHADOW, counts + Q_SHADOW_REF, counts + Q_NAN};
    flush_counts_cta<4>(v, d);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Split shade pipeline (default): k_shade_surface above cut in two at the point where the walk and the NEE part ways.
//   k_shade_vertex<CLASS> : gather path + hit, rebuild the hit, sample the BSDF, russian roulette, light-hit MIS,
//                           write the next path record; a vertex that wants NEE leaves a 64-byte hand-over record
//   k_nee<CLASS>          : reads the hand-over records (compact, sequential), draws the L light / environment samples,
//                           evaluates the BSDF and the MIS weight, appends the shadow rays
// Why (ncu, profiles/r01_final_ncu_kernels.csv + r02 captures): the fused kernel is 6 640 SASS instructions of which a
// vertex executes ~2 750; at 72 registers it spills 116 B, reaches 37 % occupancy and its top stall is instruction fetch
// (the ~35 KB straight-line path of one vertex overruns the instruction cache). Each half is about half as long, fits
// 64 registers (8 CTAs / SM, 50 % occupancy) and the NEE half reads a dense queue, so none of its lanes idle on vertices
// that hit a light or produced a NaN pdf. Cost: 64 B written + 64 B read per NEE vertex.
// Hand-over record, 4 x float4:
//   r0 = vertex point.xyz, beta              r1 = shading normal.xyz, lambda
//   r2 = wi (local, towards the previous vertex).xyz, slot bits
//   r3 = diffuse: albedo, -, -, -            ggx: eta_inner, eta_outer, kappa, material index bits
// The tangent frame is rebuilt from the normal (frame_from_normal is a pure function of it), pixel / sample from the slot.
#ifndef VERTEX_MIN_BLOCKS
#define VERTEX_MIN_BLOCKS 8  // 64 registers, 50 % occupancy
#endif
template <uint32_t CLASS>
__global__ void __launch_bounds__(SHADE_THREADS, VERTEX_MIN_BLOCKS) k_shade_vertex(DevScene S, RenderCtx R, uint32_t bounce, const PathRec *__restrict__ paths,
                                                                  const HitRec *__restrict__ hits, const uint32_t *__restrict__ queue,
                                                                  uint32_t *__restrict__ counts, uint32_t *__restrict__ next_counts,
                                                                  PathRec *__restrict__ out, NeeRec *__restrict__ nee, float *__restrict__ acc) {
  // software-pipelined gather, as in k_shade_surface: [stage][float4 k][thread]
  __shared__ float4 s_stage[2][5][SHADE_THREADS];
  const uint32_t n = counts[CLASS];
  const uint32_t n_round = (n + 31u) & ~31u;
  const uint32_t L = R.light_samples;
  const uint32_t max_bounces = R.only_direct ? 1u : R.max_bounces;
  const uint32_t tid = threadIdx.x;
  auto issue = [&](int stage, uint32_t idx) {
    if (idx != RPT_NONE) {
      const float4 *rp = reinterpret_cast<const float4 *>(paths + idx);
#pragma unroll
      for (int k = 0; k < 4; ++k) cp_async16(&s_stage[stage][k][tid], rp + k);
      cp_async16(&s_stage[stage][4][tid], hits + idx);
    }
    cp_async_commit();
  };
  const uint32_t qc = chunk_size_for(n);
  WarpChunk wc_next = chunk_init(qc), wc_nee = chunk_init(qc);
  uint32_t n_next = 0, n_nee = 0, n_nan = 0;
  auto mark_next = [&](uint32_t e) { out[e].r3 = make_float4(__uint_as_float(RPT_NONE), 0.0f, 0.0f, 0.0f); };
  auto mark_nee = [&](uint32_t e) { nee[e].r2 = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(RPT_NONE)); };
  uint32_t *nee_counter = counts + (CLASS == Q_DIFFUSE ? Q_NEE_DIFFUSE : Q_NEE_GGX);
  const uint32_t lane = tid & 31u;
  const uint32_t n_tiles = n_round >> 5;
  TileStream ts;
  ts.init(counts + (CLASS == Q_DIFFUSE ? F_SHADE_DIFFUSE : F_SHADE_GGX), n_tiles, gridDim.x * (SHADE_THREADS / 32), S.min_grab);
  auto tile_idx = [&](uint32_t t) -> uint32_t {
    uint32_t i = t * 32u + lane;
    return (t != RPT_NONE && i < n) ? __ldg(queue + i) : RPT_NONE;
  };
  uint32_t t_cur = ts.next(), t_next = ts.next();
  uint32_t idx_cur = tile_idx(t_cur), idx_next = tile_idx(t_next);
  issue(0, idx_cur);
  int stage = 0;
  while (t_cur != RPT_NONE) {
    issue(stage ^ 1, idx_next);  // prefetch the next item while this one is shaded
    uint32_t t_nn = ts.next();
    uint32_t idx_nn = tile_idx(t_nn);
    cp_async_wait<1>();  // everything but the group just committed has landed
    bool active = idx_cur != RPT_NONE;  // beyond the queue, or chunk padding
    // Three phases, so that no output record has to sit in registers across the warp-wide queue reservations:
    // (1) evaluate the vertex and decide what it emits, (2) reserve the queue entries (warp-uniform), (3) build each
    // record and store it straight to its entry.
    bool continues = false, do_nee = false;
    float3 hp = f3(0, 0, 0), hn = f3(0, 0, 1), wo = f3(0, 0, 1), prev_p = f3(0, 0, 0);
    float beta = 0.0f, lambda = 0.0f, pdf_forward = 0.0f, nbeta = 0.0f;
    uint32_t slot = 0;
    float4 extra = make_float4(0, 0, 0, 0);
    if (active) {
      PathRec r;
      r.r0 = s_stage[stage][0][tid];
      r.r1 = s_stage[stage][1][tid];
      r.r2 = s_stage[stage][2][tid];
      r.r3 = s_stage[stage][3][tid];
      float4 hraw = s_stage[stage][4][tid];
      TraceHit th;
      th.t = hraw.x;
      th.inst = __float_as_uint(hraw.y);
      th.prim = __float_as_uint(hraw.z);
      prev_p = f3(r.r0);
      float3 prev_n = f3(r.r1), d = f3(r.r2);
      float prev_pdf = r.r1.w;
      beta = r.r0.w;
      lambda = r.r2.w;
      slot = __float_as_uint(r.r3.x);
      float3 o = rec_origin(r);
      // the material id travels in the hit record (k_trace classified the vertex by it): its table entries are fetched
      // here, next to the instance / triangle loads of reconstruct_hit, instead of after them
      const uint32_t hit_mat = __float_as_uint(hraw.w);
      const RptMaterial m = S.materials[RPT_MAT_INDEX(hit_mat)];
      const float2 m_fast = __ldg(S.mat_fast + RPT_MAT_INDEX(hit_mat));
      SurfaceHit sh;
      reconstruct_hit(S, o, d, th, sh);
      sh.material = hit_mat;  // (the same id reconstruct_hit derives: instance override, else the triangle's)
      hp = sh.p;
      hn = sh.n;
      Frame frame = frame_from_normal(sh.n);
      float3 wi = normalized(to_local(frame, -d));  // integrator/utils.rs:175-176
      const uint32_t pixel = slot_to_pixel(R, slot % R.wh), sample = R.sample_base + slot / R.wh;
      RptRand4 s = rpt_philox(R.seed, pixel, sample, rpt_block_bsdf(bounce, L));

      // ---- generate_and_evaluate
      float f, pdf;
      if (CLASS == Q_GGX) {
        GgxParams gp = ggx_params(S, m, lambda);
        Bsdf b = ggx_generate_and_evaluate(gp, s.x, s.y, wi, wo);
        f = b.f;
        pdf = b.pdf;
        extra = make_float4(gp.eta_inner, gp.eta_outer, gp.kappa, __uint_as_float(RPT_MAT_INDEX(sh.material)));
      } else {
        float albedo = diffuse_albedo_fast(S, m, m_fast, lambda, sh.u, sh.v);
        wo = random_cosine_direction(s.x, s.y) * signumf(wi.z);
        f = albedo / RPT_PI;
        pdf = fabsf(wo.z) / RPT_PI;
        extra.x = albedo;
      }
      if (pdf != pdf) {
        // pdf NaN: the walk breaks BEFORE pushing the vertex (integrator/utils.rs:261-263)
        n_nan++;
      } else {
        // ---- the vertex exists: its contribution (second loop of pt.rs:481-613)
        if (RPT_MAT_IS_LIGHT(sh.material)) {
          float emission = material_emission(S, m, lambda, wi);
          if (emission > 0.0f) {
            float c = 0.0f;
            if (L == 0 || bounce == 0) {
              c = beta * emission;
            } else if (!R.only_direct) {
              float3 nee_direction = normalized(sh.p - prev_p);
              float hyp = instance_psa_pdf(S.instances[th.inst], dot(prev_n, nee_direction), dot(sh.n, nee_direction), prev_p, sh.p);
              c = power_heuristic(prev_pdf, hyp) * beta * emission;
            }
            if (c != 0.0f) atomicAdd(acc + slot, c);
          }
        } else if (L > 0) {
          do_nee = true;
        }
        // ---- continue the walk (integrator/utils.rs:266-329)
        float cos_o = fabsf(wo.z);
        float rr = bounce >= R.min_bounces ? fminf(f / pdf, 1.0f) : 1.0f;
        pdf_forward = pdf * (rr / cos_o);
        nbeta = beta * (f / pdf_forward);
        if (pdf_forward == 0.0f) nbeta = 0.0f;
        continues = nbeta != 0.0f && !(s.z > rr) && bounce + 1 < max_bounces;
      }
    }
    const uint32_t k_next = chunk_append(next_counts + Q_PATHS, wc_next, continues, mark_next, qc);
    const uint32_t k_nee = chunk_append(nee_counter, wc_nee, do_nee, mark_nee, qc);
    n_next += continues;
    n_nee += do_nee;
    if (continues || do_nee) {
      const Frame frame = frame_from_normal(hn);  // (a pure function of the normal: the same frame as above)
      if (continues) {
        float3 nd = normalized(to_world(frame, wo));
        float4 *o4 = reinterpret_cast<float4 *>(out + k_next);
        o4[0] = make_float4(hp.x, hp.y, hp.z, nbeta);
        o4[1] = make_float4(hn.x, hn.y, hn.z, pdf_forward);
        o4[2] = make_float4(nd.x, nd.y, nd.z, lambda);
        o4[3] = make_float4(__uint_as_float(slot), signumf(wo.z), 0.0f, 0.0f);
      }
      if (do_nee) {
        float3 wi_nee = to_local(frame, normalized(prev_p - hp));  // pt.rs:565-569
        float4 *n4 = reinterpret_cast<float4 *>(nee + k_nee);
        n4[0] = make_float4(hp.x, hp.y, hp.z, beta);
        n4[1] = make_float4(hn.x, hn.y, hn.z, lambda);
        n4[2] = make_float4(wi_nee.x, wi_nee.y, wi_nee.z, __uint_as_float(slot));
        n4[3] = extra;
      }
    }
    stage ^= 1;
    idx_cur = idx_next;
    idx_next = idx_nn;
    t_cur = t_next;
    t_next = t_nn;
  }
  cp_async_wait<0>();
  if (wc_next.used < qc) chunk_pad(wc_next, mark_next, qc);
  if (wc_nee.used < qc) chunk_pad(wc_nee, mark_nee, qc);
  {
    const uint32_t v[3] = {n_next, n_nee, n_nan};
    uint32_t *const d[3] = {next_counts + N_PATHS, counts + N_NEE, counts + Q_NAN};
    flush_counts_cta<3>(v, d);
  }
}

// NEE sample generation for the vertices k_shade_vertex handed over (estimate_direct_illumination_with_loop, pt.rs:333-393;
// light samples pt.rs:146-219, environment samples pt.rs:224-331). One vertex per lane, L samples each; every sample that
// can contribute becomes a shadow ray, appended binned by origin cell (see k_shade_surface).
// NEE rays are binned by the cell of the scene bounds their origin lies in (NEE_GRID^3 cells), one queue chunk of NEE_CHUNK
// rays per (warp, cell): a consumer warp of k_shadow then traces rays that start close together and (light samples) head for
// the same light.
#ifndef NEE_GRID
#define NEE_GRID 2u
#endif
#ifndef NEE_CHUNK
#define NEE_CHUNK 128u
#endif
#define NEE_BINS (NEE_GRID * NEE_GRID * NEE_GRID)
// KIND (scenes that sample BOTH the environment and lights, 0 < p_env < 1): which of the two a sample draws is a per-sample
// coin flip (pt.rs:350-353), so a warp that walks "lane = vertex, loop over its L samples" (NEE_BOTH) runs the environment
// branch (two CDF searches + the f64 uv <-> direction round trips) and the light branch with half its lanes each: ncu shows
// 15.9 lanes per instruction on the GGX + HDR scene. There the queue is walked by TWO launches, NEE_LIGHT then NEE_ENV: each
// classifies every (vertex, sample) pair (one Philox draw), parks the pairs of ITS kind in a per-warp ring buffer, and draws
// them 32 at a time, one per lane, all lanes in the same branch. An item re-reads its 64-byte vertex record (L1 / L2 hit) and
// repeats the Philox draw. One kind per launch rather than two rings in one launch: the single-launch form does reach 27-29
// lanes, but its warps sit in different halves of 90 KB of code and wait for instructions (no-instruction stalls 5-11 warps
// per issue, profiles/r02_nee_sort_*); a launch that compiles only one branch halves the footprint.
// Samples are the same samples; only the order of the shadow records (and of the energy atomics behind them) changes.
enum : int { NEE_BOTH = 0, NEE_LIGHT = 1, NEE_ENV = 2 };
enum : int { NEE_MODE_BOTH = 0, NEE_MODE_LIGHT_ONLY = 1, NEE_MODE_SORTED = 2 };
#define NEE_RING 64u
struct NeeVertex {
  float3 p, nrm, wi;
  float beta, lambda, albedo;
  uint32_t slot, pixel, sample;
  Frame frame;
  GgxParams gp;
};
template <uint32_t CLASS, int KIND, bool SORTED>
__global__ void __launch_bounds__(SHADE_THREADS, 8) k_nee(DevScene S, RenderCtx R, uint32_t bounce, const NeeRec *__restrict__ nee,
                                                         uint32_t *__restrict__ counts, float4 *__restrict__ sh_a, float4 *__restrict__ sh_b,
                                                         uint32_t *__restrict__ sh_c) {
  __shared__ WarpChunk s_chunks[SHADE_THREADS / 32][NEE_BINS];
  static_assert(!(SORTED && KIND == NEE_BOTH), "a sorted launch draws one kind");
  __shared__ uint2 s_ring[SORTED ? SHADE_THREADS / 32 : 1][SORTED ? NEE_RING : 1];  // (record, sample index) per parked item
  WarpChunk *st_shadow = s_chunks[threadIdx.x >> 5];
  const uint32_t n = counts[CLASS == Q_DIFFUSE ? Q_NEE_DIFFUSE : Q_NEE_GGX];
  const uint32_t bc = n >= BIN_MIN_ITEMS ? NEE_CHUNK : QCHUNK_BINNED_SMALL;
  for (uint32_t b = threadIdx.x & 31u; b < NEE_BINS; b += 32u) st_shadow[b] = WarpChunk{0u, bc};
  __syncwarp();
  const uint32_t L = R.light_samples;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_tiles = (n + 31u) >> 5;
  const uint32_t nbins = n >= BIN_MIN_ITEMS ? NEE_BINS : 1u;  // small queues: binning would only scatter a few rays over many chunks
  uint32_t n_shadow = 0, n_sh_ref = 0;
  auto mark_shadow = [&](uint32_t e) { sh_c[e] = RPT_NONE; };
  const float inv_l = 1.0f / (float)L;

  // the hand-over record of vertex i; returns false for chunk padding (and for lanes without a record)
  auto load_vertex = [&](uint32_t i, bool act, NeeVertex &v) -> bool {
    float4 r0 = make_float4(0, 0, 0, 0), r1 = make_float4(0, 0, 1, 0), r2 = make_float4(0, 0, 1, 0), r3 = make_float4(0, 0, 0, 0);
    if (act) {
      const float4 *rp = reinterpret_cast<const float4 *>(nee + i);
      r2 = __ldg(rp + 2);
      act = __float_as_uint(r2.w) != RPT_NONE;  // chunk padding
      if (act) {
        r0 = __ldg(rp);
        r1 = __ldg(rp + 1);
        r3 = __ldg(rp + 3);
      }
    }
    v.p = f3(r0);
    v.nrm = f3(r1);
    v.wi = f3(r2);
    v.beta = r0.w;
    v.lambda = r1.w;
    v.slot = __float_as_uint(r2.w);
    v.pixel = slot_to_pixel(R, v.slot % R.wh);
    v.sample = R.sample_base + v.slot / R.wh;
    v.frame = frame_from_normal(v.nrm);
    v.gp.alpha = 1.0f;
    v.gp.eta_inner = r3.x;
    v.gp.eta_outer = r3.y;
    v.gp.kappa = r3.z;
    v.gp.metallic = false;
    if (CLASS == Q_GGX && act) {
      const RptMaterial &m = S.materials[__float_as_uint(r3.w)];
      v.gp.alpha = m.alpha;
      v.gp.metallic = m.metallic != 0;
    }
    v.albedo = r3.x;
    return act;
  };
  // sample ls of vertex v -> at most one shadow record. All 32 lanes must call (the append is warp-collective).
  // PAIRS (the light-only loop): the entries of a vertex's samples are reserved up front, two at a time, next to each other, so
  // that neighbouring lanes of k_shadow trace rays that leave the same point; a sample that yields no ray leaves a skip marker.
#ifdef RPT_NEE_PAIRS
  constexpr bool PAIRS = !SORTED && KIND == NEE_LIGHT;
#else
  constexpr bool PAIRS = false;
#endif
  auto origin_cell = [&](float3 o) -> uint32_t {
    const float g = (float)NEE_GRID;
    uint32_t cx = (uint32_t)fminf(fmaxf((o.x - S.world_min.x) * S.world_inv_extent.x * g, 0.0f), g - 1.0f);
    uint32_t cy = (uint32_t)fminf(fmaxf((o.y - S.world_min.y) * S.world_inv_extent.y * g, 0.0f), g - 1.0f);
    uint32_t cz = (uint32_t)fminf(fmaxf((o.z - S.world_min.z) * S.world_inv_extent.z * g, 0.0f), g - 1.0f);
    return (cz * NEE_GRID + cy) * NEE_GRID + cx;
  };
  auto draw_sample = [&](const NeeVertex &v, bool do_nee, uint32_t ls, uint32_t reserved) {
    bool has = false;
    float4 a = make_float4(0, 0, 0, 0), b4 = make_float4(0, 0, 0, 0);
    uint32_t c = 0;
    if (do_nee) {
      RptRand4 sn = rpt_philox(R.seed, v.pixel, v.sample, rpt_block_nee(bounce, L, ls));
      float pick;
      bool sample_world = choose(sn.x, S.p_env, pick);  // pt.rs:350-353
      // one kind per launch: what the classification found (SORTED), or the only kind the scene has (p_env = 0, where
      // choose() cannot return anything else); lets the compiler drop the other branch
      if (KIND == NEE_LIGHT) sample_world = false;
      if (KIND == NEE_ENV) sample_world = true;
      float3 dir = f3(0, 0, 1);
      float light_pdf = 0.0f, emission = 1.0f;
      bool valid = true;
      if (sample_world) {
        float eu, ev;
        env_sample_uv(S, sn.y, sn.z, eu, ev, light_pdf);
        dir = uv_to_direction(eu, ev);
        emission = env_emission(S, eu, ev, v.lambda);
      } else {
        valid = S.num_lights > 0;
        if (valid) {
          uint32_t li = (uint32_t)clampf((float)S.num_lights * pick, 0.0f, (float)S.num_lights - 1.0f);  // world/mod.rs:109
          instance_sample(S.instances[S.lights[li]], sn.y, sn.z, v.p, dir, light_pdf);
          light_pdf = light_pdf * (1.0f / (float)S.num_lights);
          valid = light_pdf != 0.0f;  // pt.rs:151-153
        }
      }
      float3 local_wo = to_local(v.frame, dir);
      if (sample_world && local_wo.z <= 0.0f) valid = false;  // pt.rs:245-247
      if (valid) {
        Bsdf bs;
        if (CLASS == Q_GGX) {
          bs = ggx_bsdf(v.gp, v.wi, local_wo);
        } else {  // lambertian.rs:16-32 / diffuse_light.rs:29-45
          bool same = local_wo.z * v.wi.z > 0.0f;
          bs.f = same ? v.albedo / RPT_PI : 0.0f;
          bs.pdf = same ? fabsf(local_wo.z) / RPT_PI : 0.0f;
        }
        n_sh_ref++;  // the reference traces (and counts) this ray whatever its weight
        float weight = R.only_direct ? 1.0f : power_heuristic_generic(light_pdf, bs.pdf);
        float pre;
        float3 so;
        if (sample_world) {
          pre = v.beta * weight * bs.f * emission * fabsf(local_wo.z) * (1.0f / light_pdf) * inv_l;  // pt.rs:313-318
          so = v.p + (v.nrm * RPT_NORMAL_OFFSET) * signumf(dir.z);  // WORLD z (quirk Q12, pt.rs:256)
          c = v.slot | 0x80000000u;
        } else {
          pre = bs.f * v.beta * fabsf(local_wo.z) * weight / light_pdf * inv_l;  // pt.rs:196-202 minus the light-side terms
          so = v.p + (v.nrm * RPT_NORMAL_OFFSET) * signumf(local_wo.z);  // pt.rs:171-174
          c = v.slot;
        }
        if (pre != 0.0f) {  // a zero pre-factor cannot contribute: skip the visibility query
          has = true;
          a = make_float4(so.x, so.y, so.z, pre);
          b4 = make_float4(dir.x, dir.y, dir.z, v.lambda);
        }
      }
    }
    uint32_t q = reserved;
    if constexpr (!PAIRS) {
      const uint32_t bin_sh = nbins > 1 ? origin_cell(f3(a)) : 0u;
      q = chunk_append_binned(counts + Q_SHADOW, st_shadow, has, bin_sh, mark_shadow, bc);
    }
    if (has) {
      sh_a[q] = a;
      sh_b[q] = b4;
      sh_c[q] = c;
    } else if (PAIRS && do_nee) {
      sh_c[q] = RPT_NONE;
    }
    n_shadow += has;
  };

  TileStream ts;
  ts.init(counts + (KIND == NEE_ENV ? (CLASS == Q_DIFFUSE ? F_NEE_DIFFUSE_ENV : F_NEE_GGX_ENV) : (CLASS == Q_DIFFUSE ? F_NEE_DIFFUSE : F_NEE_GGX)), n_tiles,
          gridDim.x * (SHADE_THREADS / 32), S.min_grab);
  if constexpr (!SORTED) {
    for (uint32_t tile = ts.next(); tile != RPT_NONE; tile = ts.next()) {
      const uint32_t i = tile * 32u + lane;
      NeeVertex v;
      const bool do_nee = load_vertex(i, i < n, v);
      if constexpr (PAIRS) {
        const uint32_t bin_sh = nbins > 1 ? origin_cell(v.p) : 0u;
        for (uint32_t ls = 0; ls < L; ls += 2u) {
          const uint32_t per = min(2u, L - ls);
          const uint32_t base = chunk_append_binned(counts + Q_SHADOW, st_shadow, do_nee, bin_sh, mark_shadow, bc, per);
#pragma unroll 1
          for (uint32_t k = 0; k < per; ++k) draw_sample(v, do_nee, ls + k, base + k);
        }
      } else {
        for (uint32_t ls = 0; ls < L; ++ls) draw_sample(v, do_nee, ls, RPT_NONE);
      }
    }
  } else {
    uint2 *ring = s_ring[threadIdx.x >> 5];
    uint32_t head = 0u, tail = 0u;  // warp-uniform, free-running; entries live at index & (NEE_RING - 1)
    const uint32_t lt_mask = (1u << lane) - 1u;
    // producer state: the tile being classified, one sample index per step
    uint32_t i = 0u, ls = L, pixel = 0u, sample = 0u;
    bool do_nee = false, more = true;
    while (true) {
      if (tail - head < 32u && more) {
        // classify sample `ls` of the current tile's vertices (one Philox draw each) and park the pairs of this launch's kind
        if (ls == L) {
          const uint32_t tile = ts.next();
          if (tile == RPT_NONE) {
            more = false;
            continue;
          }
          i = tile * 32u + lane;
          uint32_t slot = RPT_NONE;
          if (i < n) slot = __float_as_uint(__ldg(reinterpret_cast<const float4 *>(nee + i) + 2).w);
          do_nee = slot != RPT_NONE;
          pixel = slot_to_pixel(R, slot % R.wh);
          sample = R.sample_base + slot / R.wh;
          ls = 0u;
        }
        bool mine = false;
        if (do_nee) {
          RptRand4 sn = rpt_philox(R.seed, pixel, sample, rpt_block_nee(bounce, L, ls));
          float pick;
          mine = choose(sn.x, S.p_env, pick) == (KIND == NEE_ENV);
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
        if (mine) ring[(tail + __popc(m & lt_mask)) & (NEE_RING - 1u)] = make_uint2(i, ls);
        tail += __popc(m);
        ++ls;
        __syncwarp();
        continue;
      }
      if (tail == head) break;
      // up to 32 parked items, one per lane, every lane in the same branch of draw_sample
      const uint32_t cnt = min(32u, tail - head);
      const bool act = lane < cnt;
      uint2 item = make_uint2(0u, 0u);
      if (act) item = ring[(head + lane) & (NEE_RING - 1u)];
      head += cnt;
      __syncwarp();
      NeeVertex v;
      const bool ok = load_vertex(item.x, act, v);
      draw_sample(v, ok, item.y, RPT_NONE);
    }
  }
  chunk_pad_binned(st_shadow, NEE_BINS, mark_shadow, bc);
  {
    const uint32_t v[2] = {n_shadow, n_sh_ref};  // Q_SHADOW_REF: reference-definition shadow-ray counter (pt.rs:176,252)
    uint32_t *const d[2] = {counts + N_SHADOW, counts + Q_SHADOW_REF};
    flush_counts_cta<2>(v, d);
  }
}

// Phase A of the two-phase NEE visibility query: the closest hit among the scene's few analytic light-material shapes
// (DevScene::light_geom), with the reference's tie rule. Returns false when the ray meets none of them.
template <bool STATS>
__device__ __forceinline__ bool shadow_closest_light(const DevScene &S, float3 o, float3 d, float &tl, uint64_t &key_l, uint32_t &inst_l, TraceWork &tw) {
  bool found_l = false;
  for (uint32_t k = 0; k < S.num_light_geom; ++k) {
    uint32_t li = __ldg(S.light_geom + k);
    const DevInstance &I = S.instances[li];
    uint32_t flags = I.flags;
    if ((flags & DI_KIND_MASK) == RPT_AGG_DISK) {
      // A disk's bounding box is SMALLER than the disk (disk.rs:23-28: half extent radius / 2), so the part of the
      // disk outside it is unreachable through the reference's BVH; the traversal reproduces that through the leaf
      // box, phase A has to gate on the same box.
      const float *bx = S.light_geom_box + 6 * k;
      float3 winv, woinv;
      slab_recip(o, d, winv, woinv);
      float tn;
      if (!slab_test(f3(__ldg(bx), __ldg(bx + 1), __ldg(bx + 2)), f3(__ldg(bx + 3), __ldg(bx + 4), __ldg(bx + 5)), woinv, winv, tl, tn)) continue;
    }
    float3 lo = o, ld = d;
    if (flags & DI_HAS_TRANSFORM) {
      lo = xform_point(I.rev, o);
      ld = xform_vec(I.rev, d);
    }
    uint32_t kind = flags & DI_KIND_MASK;
    float t;
    bool hit = kind == RPT_AGG_RECT ? rect_test(I, lo, ld, 0.0f, tl, RPT_INF, t)
                                    : (kind == RPT_AGG_SPHERE ? sphere_test(I, lo, ld, 0.0f, tl, RPT_INF, t) : disk_test(I, lo, ld, 0.0f, tl, RPT_INF, t));
    RPT_STAT(tw.insts++);
    if (hit) {
      uint64_t key = tie_key(kind == RPT_AGG_SPHERE, I.order, 0);
      if (!found_l || t < tl || key > key_l) {
        tl = t;
        key_l = key;
        found_l = true;
        inst_l = li;
      }
    }
  }
  return found_l;
}
// The light sample's contribution once its ray is known to end on a light-material surface (pt.rs:191-218): the emission
// of THAT surface towards the ray, times the light-side cosine (quirk Q14).
__device__ __forceinline__ void shadow_light_contribution(const DevScene &S, float3 o, float3 d, const TraceHit &th, float pre, float lambda, uint32_t slot,
                                                          float *__restrict__ acc) {
  SurfaceHit sh;
  reconstruct_hit(S, o, d, th, sh);
  if (RPT_MAT_IS_LIGHT(sh.material)) {
    Frame lf = frame_from_normal(sh.n);
    float3 lwi = to_local(lf, -d);
    float le = material_emission(S, S.materials[RPT_MAT_INDEX(sh.material)], lambda, lwi);
    float v = pre * fabsf(lwi.z) * le;
#ifdef RPT_DEBUG
    if (slot == RPT_DEBUG_SLOT) printf("[shadow] hit inst=%u prim=%u t=%.7g le=%.7g lwi.z=%.7g pre=%.7g -> %.7g\n", th.inst, th.prim, th.t, le, lwi.z, pre, v);
#endif
    if (v != 0.0f) atomicAdd(acc + slot, v);
  }
}

// NEE visibility. Light samples: closest hit, accepted when ANY light-material surface is hit, whose own
// emission is used (pt.rs:177-218, F9). Environment samples: any hit kills the sample (pt.rs:254-263).
template <int MODE, bool STATS>
__global__ void __launch_bounds__(TRACE_THREADS, MODE == 4 ? TRACE_MIN_BLOCKS_WIDE : (MODE == 5 ? TRACE_MIN_BLOCKS_FLAT : TRACE_MIN_BLOCKS)) k_shadow(DevScene S, const float4 *__restrict__ sh_a, const float4 *__restrict__ sh_b,
                                                          const uint32_t *__restrict__ sh_c, uint32_t *__restrict__ counts,
                                                          float *__restrict__ acc, unsigned long long *__restrict__ work) {
  extern __shared__ int s_stack[];
  constexpr bool SMALL = MODE == TRAV_SMALL;
  constexpr bool WIDE = MODE == TRAV_BVH4;
  constexpr bool TWO_LEVEL = MODE != TRAV_BVH_FLAT;
  const float4 *s_tris = reinterpret_cast<const float4 *>(s_stack);
  if (SMALL) stage_small_tris(S, reinterpret_cast<float4 *>(s_stack));
  TraceWork tw{0, 0, 0};
  const uint32_t n = counts[Q_SHADOW];
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_tiles = (n + 31u) >> 5;
  // SMALL: every lane of the warp walks the leaf list together (SmallTrav), so lanes without a ray stay in the body with
  // active == false instead of returning early.
  auto body = [&](uint32_t i) {
    bool active = i < n;
    uint32_t c = active ? __ldg(sh_c + i) : RPT_NONE;
    active = active && c != RPT_NONE;  // chunk padding
    if (!SMALL && !active) return;
    float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 1, 0);
    if (active) {
      a = __ldg(sh_a + i);
      b = __ldg(sh_b + i);
    }
    float3 o = f3(a), d = f3(b);
    float pre = a.w, lambda = b.w;
    uint32_t slot = c & 0x7FFFFFFFu;
    TraceHit th;
    const bool env_ray = active && (c & 0x80000000u);
    if (SMALL) {
      // environment samples (any hit kills the sample) and light samples share one walk of the leaf list: a light
      // sample first finds the closest light-material shape (phase A below), then both kinds look for ANY leaf that
      // beats what they hold (nothing / that light).
      SmallTrav sv;
      sv.init(RPT_INF);
      bool search = active;
      uint64_t key_l = 0;
      float tl = RPT_INF;
      if (S.num_light_geom) {
        if (active && !env_ray) {
          uint32_t inst_l = RPT_NONE;
          bool found_l = shadow_closest_light<STATS>(S, o, d, tl, key_l, inst_l, tw);
          search = found_l;  // no light along the ray: nothing to add
          sv.closest = tl;
          sv.best_key = key_l;
          sv.found = found_l;
          sv.out.t = tl;
          sv.out.inst = inst_l;
          sv.out.prim = 0;
        }
        sv.template run<true, STATS>(S, s_tris, o, d, RPT_INF, search, tw);
      } else {
        // light-material geometry is not a short analytic list: closest hit for light samples, any hit for environment
        // samples; the two kinds are walked separately so that each loop stays warp-uniform
        sv.template run<true, STATS>(S, s_tris, o, d, RPT_INF, env_ray, tw);
        SmallTrav sl;
        sl.init(RPT_INF);
        sl.template run<false, STATS>(S, s_tris, o, d, RPT_INF, active && !env_ray, tw);
        if (active && !env_ray) sv = sl;
      }
      if (!active) return;
      if (env_ray) {
        if (!sv.found) atomicAdd(acc + slot, pre);
        return;
      }
      bool lit = S.num_light_geom ? (search && sv.best_key == key_l && sv.closest == tl) : sv.found;
      if (lit) shadow_light_contribution(S, o, d, sv.out, pre, lambda, slot, acc);
      return;
    }
    if (env_ray) {
      if (!trace_ray<true, STATS, WIDE, TWO_LEVEL>(S, o, d, RPT_INF, s_stack + threadIdx.x, TRACE_THREADS, th, tw)) atomicAdd(acc + slot, pre);
      return;
    }
    bool lit;
    if (S.num_light_geom) {
      // Two-phase form of "the closest hit must be a light" (exact, including the reference's tie rule):
      // (A) the closest hit among the few light-material shapes, (B) ANY hit of the scene in front of it.
      float tl = RPT_INF;
      uint64_t key_l = 0;
      uint32_t inst_l = RPT_NONE;
      if (!shadow_closest_light<STATS>(S, o, d, tl, key_l, inst_l, tw)) return;  // no light along the ray: nothing to add
      TravT<WIDE, TWO_LEVEL> tv;
      tv.init(S, o, d, RPT_INF);
      tv.closest = tl;  // only geometry that beats the light (closer, or equal t with a winning tie key) is accepted
      tv.best_key = key_l;
      tv.found = true;
      tv.template run<true, STATS>(S, s_stack + threadIdx.x, TRACE_THREADS, tw);
      lit = tv.best_key == key_l && tv.closest == tl;
      th.t = tl;
      th.inst = inst_l;
      th.prim = 0;
    } else {
      lit = trace_ray<false, STATS, WIDE, TWO_LEVEL>(S, o, d, RPT_INF, s_stack + threadIdx.x, TRACE_THREADS, th, tw);
    }
    if (lit) shadow_light_contribution(S, o, d, th, pre, lambda, slot, acc);
  };
  if (MODE == TRAV_BVH_REFILL) {
    // Lane refill (see TRAV_BVH_REFILL above). A lane keeps only its queue index and the phase-A result; the record's
    // pre-factor / wavelength / slot are re-read when the ray is finished. Both kinds of ray walk with the closest-hit step;
    // an environment ray is finished by its first accepted hit, a two-phase light ray by the first hit that beats its light.
    TileStream ts;
    ts.init(counts + F_SHADOW, n_tiles, gridDim.x * (TRACE_THREADS / 32), S.min_grab);
    uint32_t pool_cur = 0, pool_end = 0, my_i = 0, inst_l = RPT_NONE;
    bool has = false, exhausted = false, env = false;
    float tl = RPT_INF;
    uint64_t key_l = 0;
    Trav t;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool two_phase = S.num_light_geom != 0;
    while (true) {
      const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !has);
      if (!exhausted && (__popc(idle) >= REFILL_MIN || idle == 0xFFFFFFFFu)) {
        const uint32_t need = __popc(idle), rank = __popc(idle & lt_mask);
        uint32_t given = 0;
        while (given < need) {
          if (pool_cur == pool_end) {
            const uint32_t tile = ts.next();
            if (tile == RPT_NONE) {
              exhausted = true;
              break;
            }
            pool_cur = tile * 32u;
            pool_end = min(pool_cur + 32u, n);
          }
          const uint32_t take = min(need - given, pool_end - pool_cur);
          if (!has && rank >= given && rank < given + take) {
            my_i = pool_cur + (rank - given);
            const uint32_t c = __ldg(sh_c + my_i);
            if (c != RPT_NONE) {  // (chunk padding: the lane stays idle until the next hand-out)
              const float3 o = f3(__ldg(sh_a + my_i)), d = f3(__ldg(sh_b + my_i));
              env = (c & 0x80000000u) != 0;
              bool go = true;
              t.init(S, o, d, RPT_INF);
              if (!env && two_phase) {
                tl = RPT_INF;
                key_l = 0;
                inst_l = RPT_NONE;
                go = shadow_closest_light<STATS>(S, o, d, tl, key_l, inst_l, tw);  // no light along the ray: nothing to add
                t.closest = tl;
                t.best_key = key_l;
                t.found = true;
              }
              has = go;
            }
          }
          pool_cur += take;
          given += take;
        }
      }
      if (!__any_sync(0xFFFFFFFFu, has)) {
        if (exhausted) break;
        continue;
      }
      if (has) {
        bool finished = t.template step<false, STATS>(S, s_stack + threadIdx.x, TRACE_THREADS, tw);
        const bool beaten = env ? t.found : (two_phase && (t.best_key != key_l || t.closest != tl));
        if (finished || beaten) {
          has = false;
          const float pre = __ldg(sh_a + my_i).w, lambda = __ldg(sh_b + my_i).w;
          const uint32_t slot = __ldg(sh_c + my_i) & 0x7FFFFFFFu;
          if (env) {
            if (!t.found) atomicAdd(acc + slot, pre);
          } else {
            TraceHit th = t.out;
            bool lit = t.found;
            if (two_phase) {
              lit = !beaten;
              th.t = tl;
              th.inst = inst_l;
              th.prim = 0;
            }
            if (lit) shadow_light_contribution(S, t.o, t.d, th, pre, lambda, slot, acc);
          }
        }
      }
    }
  } else if (!TRACE_DYNAMIC && !SMALL) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) body(i);
  } else {
    TileStream ts;
    ts.init(counts + F_SHADOW, n_tiles, gridDim.x * (TRACE_THREADS / 32), S.min_grab);
    for (uint32_t tile = ts.next(); tile != RPT_NONE; tile = ts.next()) body(tile * 32u + lane);
  }
  if (STATS) flush_work(tw, work);
}

// XYZColor::from(SingleWavelength) (pt.rs:614) + per-pixel accumulation (tiled.rs:390).
// One thread per pixel sums that pixel's samples of the wave: coalesced, no atomics. The CIE tables are
// staged in shared memory once per CTA.
__global__ void __launch_bounds__(256) k_film(DevScene S, RenderCtx R, const float *__restrict__ acc, float4 *__restrict__ film) {
  extern __shared__ float s_cie[];
  for (uint32_t i = threadIdx.x; i < 3 * S.num_lambda; i += blockDim.x) s_cie[i] = S.cie_lut[i];
  __syncthreads();
  uint32_t spp = R.n_slots / R.wh;
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < R.wh; q += stride) {
    const uint32_t pixel = slot_to_pixel(R, q);
    float X = 0.0f, Y = 0.0f, Z = 0.0f;
    for (uint32_t s = 0; s < spp; ++s) {
      float e = acc[(size_t)s * R.wh + q];
      if (e == 0.0f) continue;
      RptRand4 s0 = rpt_philox(R.seed, pixel, R.sample_base + s, 0);
      float lambda = R.lambda_lo + s0.z * (R.lambda_hi - R.lambda_lo);
      float x = (lambda - S.lut_lo) / (S.lut_hi - S.lut_lo) * (float)(S.num_lambda - 1);
      x = clampf(x, 0.0f, (float)(S.num_lambda - 1));
      uint32_t j = (uint32_t)x;
      if (j > S.num_lambda - 2) j = S.num_lambda - 2;
      float t = x - (float)j;
      const float *cx = s_cie, *cy = s_cie + S.num_lambda, *cz = s_cie + 2 * S.num_lambda;
      X += e * __fadd_rn(cx[j], __fmul_rn(t, __fsub_rn(cx[j + 1], cx[j])));
      Y += e * __fadd_rn(cy[j], __fmul_rn(t, __fsub_rn(cy[j + 1], cy[j])));
      Z += e * __fadd_rn(cz[j], __fmul_rn(t, __fsub_rn(cz[j + 1], cz[j])));
    }
    float4 f = film[pixel];
    f.x += X;
    f.y += Y;
    f.z += Z;
    film[pixel] = f;
  }
}

__global__ void k_scale(float4 *film, uint64_t n, float s) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 f = film[i];
    f.x *= s;
    f.y *= s;
    f.z *= s;
    f.w = 0.0f;
    film[i] = f;
  }
}

// ---- output_film on the device: Tonemapper::initialize = one reduction over the film,
// Tonemapper::map + XYZ->RGB + OETF + byte encoding = one map kernel (renderer/mod.rs:24-80, tonemap/*.rs).
__constant__ float c_xyz_to_rec709[9] = {3.24096994f, -1.53738318f, -0.49861076f, -0.96924364f, 1.8759675f, 0.04155506f, 0.05563008f, -0.20397696f, 1.05697151f};
__constant__ float c_xyz_to_rec2020[9] = {1.4628067f, -0.1840623f, -0.2743606f, -0.5217933f, 1.4472381f, 0.0677227f, 0.0349342f, -0.0968930f, 1.2884099f};

// sums[0..3] += sum over non-NaN-luminance pixels of ln(0.001 + channel) (double accumulation: the reference sums
// sequentially in f64 (luminance variants) or f32 (x3 variants); a parallel double sum is within 1e-6 of either)
__global__ void __launch_bounds__(256) k_out_reduce(const float4 *__restrict__ film, uint64_t n, int mode /*0 y f64-delta, 1 y f32-delta, 2 x3*/,
                                                    double *__restrict__ sums) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 c = film[i];
    if (c.y != c.y) continue;
    if (mode == 2) {
      acc[0] += (double)logf(0.001f + c.x);
      acc[1] += (double)logf(0.001f + c.y);
      acc[2] += (double)logf(0.001f + c.z);
      acc[3] += (double)logf(0.001f + c.w);
    } else if (mode == 0) {
      acc[1] += log(0.001 + (double)c.y);
    } else {
      acc[1] += log((double)(0.001f + c.y));
    }
  }
  __shared__ double s_red[4][8];
  for (int k = 0; k < 4; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += s_red[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, v);
  }
}

// The x3 tonemappers accumulate ln(0.001 + XYZW) in f32, pixel by pixel (reinhard0.rs:154, reinhard1.rs:163): at
// 2 M pixels that sum carries an order-dependent rounding error of a few 1e-4 relative, which moves l_w and with it
// every output byte. To return the reference's numbers the device reproduces the ORDER: logs in parallel, then one
// strictly sequential f32 chain per channel (4 lanes, ~4 cycles per add: ~4 ms for a 1080p film, once per frame).
__global__ void __launch_bounds__(256) k_out_log(const float4 *__restrict__ film, uint64_t n, float4 *__restrict__ logs) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 c = film[i];
    logs[i] = (c.y != c.y) ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : make_float4(logf(0.001f + c.x), logf(0.001f + c.y), logf(0.001f + c.z), logf(0.001f + c.w));
  }
}
__global__ void k_out_seqsum(const float *__restrict__ logs, uint64_t n, double *__restrict__ sums) {
  if (threadIdx.x < 4) {
    float s = 0.0f;
    const float *p = logs + threadIdx.x;
#pragma unroll 8
    for (uint64_t i = 0; i < n; ++i) s = __fadd_rn(s, p[4 * i]);
    sums[threadIdx.x] = (double)s;
  }
}

// ---- N3: ImportanceMap::bake_raw (world/importance_map.rs:78-253) ---------------------------------------------
// One thread per texel of the map: the 100-sample spectral integral of luminance(lambda) * texture_stack.curve_at(uv)(lambda)
// with the Machine clamps in their reference places (see include/rpt.h). Texels come from the scene's resident
// environment textures; curves arrive pre-evaluated at the sample wavelengths.
__global__ void __launch_bounds__(256) k_imap_texel(DevScene S, uint32_t rows, uint32_t cols, uint32_t ns, float step,
                                                    const float *__restrict__ lum, const float *__restrict__ basis, float *__restrict__ texel_lum) {
  const RptTexStack st = S.stacks[S.env_texstack];
  const uint64_t n = (uint64_t)rows * cols, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    uint32_t row = (uint32_t)(idx / cols), col = (uint32_t)(idx % cols);
    float u = (float)row / (float)rows, v = (float)col / (float)cols;  // :137-140
    float uu = clampf(u, 0.0f, 1.0f - RPT_EPS), vv = clampf(v, 0.0f, 1.0f - RPT_EPS);  // vec2d.rs:34-42
    float sum = 0.0f;
    for (uint32_t i = 0; i < ns; ++i) {
      float stack = 0.0f;
      for (uint32_t k = 0; k < st.count; ++k) {
        const DevTexture &T = S.textures[S.stack_tex[st.first + k]];
        size_t x = (size_t)(uu * (float)T.width), y = (size_t)(vv * (float)T.height);
        const float *tx = T.texels + (y * T.width + x) * T.channels;
        const float *bs = basis + (size_t)(4 * k) * ns;
        float tv;
        if (T.channels == 1) {
          tv = fmaxf(__ldg(tx) * __ldg(bs + i), 0.0f);
        } else {
          tv = 0.0f;
#pragma unroll
          for (int c = 0; c < 4; ++c) tv = tv + fmaxf(__ldg(tx + c) * __ldg(bs + (size_t)c * ns + i), 0.0f);
          tv = fmaxf(tv, 0.0f);
        }
        stack = stack + tv;
      }
      stack = fmaxf(stack, 0.0f);
      sum += fmaxf(1.0f * __ldg(lum + i) * stack, 0.0f);
    }
    texel_lum[idx] = sum * step;
  }
}

// One thread per row: the reference accumulates the row mass sequentially in f32 (:153-156); a parallel scan would round
// differently, and 1024 dependent adds per row are nothing. Then the per-row normalisation (:158-163).
__global__ void __launch_bounds__(128) k_imap_rows(uint32_t rows, uint32_t cols, const float *__restrict__ texel_lum, float *__restrict__ row_pdf,
                                                   float *__restrict__ row_cdf, float *__restrict__ row_sum) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const float *src = texel_lum + (size_t)row * cols;
  float *pdf = row_pdf + (size_t)row * cols, *cdf = row_cdf + (size_t)row * cols;
  float acc = 0.0f;
  for (uint32_t c = 0; c < cols; ++c) {
    acc += src[c];
    cdf[c] = acc;
  }
  for (uint32_t c = 0; c < cols; ++c) {
    pdf[c] = src[c] / acc;
    cdf[c] = cdf[c] / acc;
  }
  row_sum[row] = acc;
}

// Marginal: row sums / total (:199,214) and Curve::Linear::to_cdf (:239-244). Sequential by definition, one thread.
__global__ void k_imap_marginal(uint32_t rows, const float *__restrict__ row_sum, float *__restrict__ m_pdf, float *__restrict__ m_cdf,
                                float *__restrict__ integral_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  float total = 0.0f;
  for (uint32_t r = 0; r < rows; ++r) total += row_sum[r];
  float mstep = (1.0f - 0.0f) / (float)rows, acc = 0.0f;
  for (uint32_t r = 0; r < rows; ++r) {
    float p = row_sum[r] / total;
    m_pdf[r] = p;
    acc += p * mstep;
    m_cdf[r] = acc;
  }
  for (uint32_t r = 0; r < rows; ++r) m_cdf[r] = m_cdf[r] / acc;
  *integral_out = acc;
}

__device__ __forceinline__ float oetf_dev(float v, uint32_t cs) {
  if (cs == RPT_COLORSPACE_SRGB) return v < 0.0031308f ? (323.0f / 25.0f) * v : (211.0f / 200.0f) * powf(v, 5.0f / 12.0f) - (11.0f / 200.0f);
  return v < 0.01805397f ? 4.5f * v : 1.0992968f * powf(v, 0.45f) - 0.09929682f;
}

__global__ void __launch_bounds__(256) k_out_map(const float4 *__restrict__ film, uint64_t n, RptOutputSettings O, float4 lw,
                                                 float *__restrict__ rgb_linear, uchar4 *__restrict__ rgba8) {
  const float *M = O.colorspace == RPT_COLORSPACE_REC2020 ? c_xyz_to_rec2020 : c_xyz_to_rec709;
  const float3 mauve = f3(0.5199467f, 51.48687f, 1.0180528f);  // src/lib.rs:46
  const bool x3 = !O.luminance_only && O.tonemapper != RPT_TONEMAP_CLAMP;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 c = film[i];
    if (rgb_linear) {  // EXR payload (tonemap/mod.rs:225-247)
      float x = O.factor * c.x, y = O.factor * c.y, z = O.factor * c.z;
      for (int r = 0; r < 3; ++r) rgb_linear[3 * i + r] = M[3 * r] * x + M[3 * r + 1] * y + M[3 * r + 2] * z;
    }
    bool fin = isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && isfinite(c.w);
    float3 m;
    if (O.tonemapper == RPT_TONEMAP_CLAMP) {  // clamp.rs:76-101
      float4 v = make_float4(c.x * O.factor, c.y * O.factor, c.z * O.factor, c.w * O.factor);
      float3 col = f3(v);
      if (!(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w))) col = mauve;
      float em = powf(2.0f, O.exposure);
      if (O.luminance_only) {
        float sf = clampf(col.y * em, 0.0f, 1.0f) / col.y;
        m = sf * col;
      } else {
        m = f3(fmaxf(fminf(col.x * em, 1.0f), 0.0f), fmaxf(fminf(col.y * em, 1.0f), 0.0f), fmaxf(fminf(col.z * em, 1.0f), 0.0f));
      }
    } else if (!x3) {  // reinhard0.rs:96-113 / reinhard1.rs:88-107
      float l = O.key_value * c.y / lw.y;
      float sf;
      if (O.tonemapper == RPT_TONEMAP_REINHARD0) {
        sf = l / (1.0f + l);
      } else {
        float mul = 1.0f / (O.white_point * O.white_point);
        sf = l * (mul * l + 1.0f) / (1.0f + l);
      }
      float3 col = fin ? f3(c) : mauve;
      m = sf * col;
    } else {  // per channel (reinhard0.rs:196-213 / reinhard1.rs:198-232)
      float3 col = (O.tonemapper == RPT_TONEMAP_REINHARD0 && !fin) ? mauve : f3(c);
      float cc[3] = {c.x, c.y, c.z}, lwv[3] = {lw.x, lw.y, lw.z}, cv[3] = {col.x, col.y, col.z}, mm[3];
      bool bad = false;
      for (int k = 0; k < 3; ++k) {
        float l = O.key_value * cc[k] / lwv[k];
        float sf;
        if (O.tonemapper == RPT_TONEMAP_REINHARD0) {
          sf = l / (1.0f + l);
        } else {
          float mul = 1.0f / powf(O.white_point, 2.0f);
          sf = l * (mul * l + 1.0f) / (1.0f + l);
        }
        mm[k] = sf * cv[k];
        bad |= !isfinite(mm[k]);
      }
      m = f3(mm[0], mm[1], mm[2]);
      if (O.tonemapper == RPT_TONEMAP_REINHARD1 && bad) m = mauve;
    }
    unsigned char b[3];
    for (int r = 0; r < 3; ++r) {
      float lin = M[3 * r] * m.x + M[3 * r + 1] * m.y + M[3 * r + 2] * m.z;
      float e = ceilf(oetf_dev(lin, O.colorspace) * 255.0f);
      e = e < 0.0f ? 0.0f : (e > 255.0f ? 255.0f : e);
      b[r] = (e != e) ? 0 : (unsigned char)e;  // NaN -> 0 like Rust's `as u8`
    }
    rgba8[i] = make_uchar4(b[0], b[1], b[2], 255);
  }
}

// ---- bandwidth probes (rpt_probe_bandwidth): the denominators of the roofline fractions, measured on the box ------------
// mode 0: streaming 128-bit reads of the whole buffer, every repetition (buffer <= L2: L2 bandwidth; >> L2: HBM read bandwidth)
// mode 1: copy first half -> second half (read + write bytes counted)
// mode 2: dependent-free random 64-byte record gathers (4 x 16 B per lane, the BVH node fetch pattern) within the buffer
__global__ void __launch_bounds__(256) k_probe_bw(const float4 *__restrict__ src, float4 *__restrict__ dst, uint64_t n16, uint32_t reps, int mode,
                                                  float *__restrict__ sink) {
  float acc = 0.0f;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  if (mode == 0) {
    for (uint32_t r = 0; r < reps; ++r)
      for (uint64_t i = tid; i < n16; i += stride) {
        float4 v;
        asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));  // L2-coherent: bypass L1
        acc += v.x + v.y + v.z + v.w;
      }
  } else if (mode == 1) {
    const uint64_t half = n16 / 2;
    for (uint32_t r = 0; r < reps; ++r)
      for (uint64_t i = tid; i < half; i += stride) dst[half + i] = src[i];
  } else {
    // every thread issues `reps` independent gathers whatever the footprint, so small (cache-resident) footprints are
    // measured at the same parallelism as large ones
    const uint64_t nrec = n16 / 4;
    uint64_t x = tid * 0x9E3779B97F4A7C15ull + 12345u;
    for (uint32_t r = 0; r < reps; ++r) {
      x = x * 6364136223846793005ull + 1442695040888963407ull;
      const float4 *p = src + 4 * ((x >> 20) % nrec);
      float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
      acc += a.x + b.y + c.z + d.w;
    }
  }
  if (acc == 123.456f) *sink = acc;  // keeps the loads alive
}

// Guide tables of the importance map's CDF inversions (DevScene::imap_*_guide): one thread per (CDF, entry).
__global__ void __launch_bounds__(256) k_imap_guides(uint32_t n_cdfs, uint32_t n, const float *__restrict__ cdfs, uint32_t *__restrict__ guides) {
  const size_t total = (size_t)n_cdfs * RPT_IMAP_GUIDE;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(t / RPT_IMAP_GUIDE), k = (uint32_t)(t % RPT_IMAP_GUIDE);
    const float *cdf = cdfs + (size_t)r * n;
    const float top = nearest_curve_eval(cdf, n, 1.0f - 0.0001f);  // what nearest_cdf_sample scales its sample by
    guides[t] = cdf_lower_bound(cdf, 0u, n, imap_guide_threshold(k, top));
  }
}

__global__ void k_debug_env_roundtrip(uint32_t n, const float *__restrict__ u, const float *__restrict__ v, int which, float *__restrict__ uo,
                                      float *__restrict__ vo) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 ident[3] = {make_float4(1, 0, 0, 0), make_float4(0, 1, 0, 0), make_float4(0, 0, 1, 0)};
  const float2 q = which ? uv_roundtrip_unrotated_cr(u[i], v[i]) : direction_to_uv_cr(xform_vec(ident, uv_to_direction_cr(u[i], v[i])));
  uo[i] = q.x;
  vo[i] = q.y;
}

// generic closest-hit query of host-provided rays (rpt_trace_rays)
template <int MODE>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_rays(DevScene S, uint32_t n, const float *__restrict__ o, const float *__restrict__ d,
                                                              const float *__restrict__ tmax, HitRec *__restrict__ hits) {
  extern __shared__ int s_stack[];
  if (MODE == TRAV_SMALL) stage_small_tris(S, reinterpret_cast<float4 *>(s_stack));
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t n_round = (n + 31u) & ~31u;  // whole warps stay in the loop (TRAV_SMALL walks the leaf list warp-wide)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool active = i < n;
    TraceHit th;
    TraceWork tw{0, 0, 0};
    float3 ro = f3(0, 0, 0), rd = f3(0, 0, 1);
    float tm = RPT_INF;
    if (active) {
      ro = f3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
      rd = f3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
      tm = tmax[i];
    }
    if (MODE == TRAV_SMALL) {
      SmallTrav sv;
      sv.init(tm);
      sv.template run<false, false>(S, reinterpret_cast<const float4 *>(s_stack), ro, rd, tm, active, tw);
      th = sv.out;
    } else if (active) {
      trace_ray<false, false, MODE == TRAV_BVH4, MODE != TRAV_BVH_FLAT>(S, ro, rd, tm, s_stack + threadIdx.x, TRACE_THREADS, th, tw);
    }
    if (active) {
      HitRec h;
      h.t = th.t;
      h.inst = th.inst;
      h.prim = th.prim;
      h.mat = RPT_NONE;
      hits[i] = h;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side: scene upload + wave driver
// ---------------------------------------------------------------------------------------------
// K_SHADE_*: k_shade_vertex<> of the split pipeline, or the fused k_shade_surface<> (RPT_FUSED_SHADE=1); K_NEE_*: k_nee<>
enum KernelId { K_RAYGEN, K_TRACE, K_SHADE_MISS, K_SHADE_DIFFUSE, K_SHADE_GGX, K_NEE_DIFFUSE, K_NEE_GGX, K_SHADOW, K_FILM, K_NUM };
const char *kKernelNamesSplit[K_NUM] = {"k_raygen", "k_trace", "k_shade_miss", "k_shade_vertex<diffuse>", "k_shade_vertex<ggx>", "k_nee<diffuse>", "k_nee<ggx>", "k_shadow", "k_film"};
const char *kKernelNamesFused[K_NUM] = {"k_raygen", "k_trace", "k_shade_miss", "k_shade_surface<diffuse>", "k_shade_surface<ggx>", "-", "-", "k_shadow", "k_film"};

// Scene arrays are sub-allocated from one device block (small scenes: a single 4 MB block that is handed from scene to
// scene through a per-device spare, so a create / render / destroy frame loop issues no cudaMalloc / cudaFree for them:
// each of those costs 0.1-8 ms and cudaFree synchronises the device). Arrays larger than 1 MB get their own allocation.
struct DeviceBuffers {
  static constexpr size_t kBlock = (size_t)4 << 20;
  std::vector<void *> ptrs;  // every cudaMalloc'd block this scene owns
  char *block = nullptr;
  size_t block_size = 0, block_used = 0;
  int device = 0;
  static void *&spare(int device) {
    static void *s[64] = {nullptr};
    return s[device & 63];
  }
  int reserve(size_t bytes, void **out) {
    *out = nullptr;
    if (bytes > ((size_t)1 << 20)) {
      CUDA_TRY(cudaMalloc(out, bytes));
      ptrs.push_back(*out);
      return 0;
    }
    size_t at = (block_used + 255) & ~(size_t)255;
    if (!block || at + bytes > block_size) {
      void *p = nullptr;
      {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        p = spare(device);
        spare(device) = nullptr;
      }
      if (!p) CUDA_TRY(cudaMalloc(&p, kBlock));
      ptrs.push_back(p);
      block = static_cast<char *>(p);
      block_size = kBlock;
      at = 0;
    }
    *out = block + at;
    block_used = at + bytes;
    return 0;
  }
  template <class T>
  int upload(const T *host, size_t n, const T **out) {
    *out = nullptr;
    if (n == 0) return 0;
    void *p = nullptr;
    if (int rc = reserve(n * sizeof(T), &p)) return rc;
    CUDA_TRY(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T *>(p);
    return 0;
  }
  void release() {
    for (void *p : ptrs) {
      if (!p) continue;
      if (p == block && block_size == kBlock) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (!spare(device)) {
          spare(device) = p;  // keep one standard block per device for the next scene
          continue;
        }
      }
      cudaFree(p);
    }
    ptrs.clear();
    block = nullptr;
    block_size = block_used = 0;
  }
};

}  // namespace

struct RptScene {
  int device = 0;
  int num_sms = 0;
  DevScene dev{};
  DeviceBuffers bufs;
  std::vector<RptCamera> cameras;
  RptSceneStats stats{};
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // the second half-wave's stream (render_waves), created on first use
  cudaEvent_t ev_join = nullptr;   // fork / join between the two
  // wave buffers (grown on demand, reused across calls)
  WaveBuffers wave{};
  size_t wave_slots = 0, wave_shadow = 0, wave_acc = 0;  // capacities: path-queue entries, shadow entries, energy slots
  float4 *film = nullptr;
  size_t film_pixels = 0;
  // launch geometry
  int grid[K_NUM] = {0};
  uint32_t stack_entries = 16;
  size_t stack_smem = 0;
  uint32_t env_stack_count = 0;  // textures in the environment's stack (HDR)
  bool has_ggx = true;     // any material of the GGX class (else its shade kernel is never launched)
  bool fused_shade = false;  // RPT_FUSED_SHADE=1: the round-1 single shade kernel instead of k_shade_vertex + k_nee
  bool flat_tlas = false;  // no transformed mesh instance: every TLAS leaf is a triangle or an analytic shape
  int nee_mode = 0;  // NEE_MODE_*: which form of k_nee the scene's NEE sample generation takes
  int trav_mode = TRAV_BVH;  // TRAV_SMALL (RPT_SMALL=1) for scenes of <= RPT_SMALL_MAX leaves without a BLAS;
                             // TRAV_BVH_TMA: k_trace reads its queue through TMA-staged shared-memory tiles (RPT_TMA_TILES=1)
  size_t counts_cap = 0;     // bounces the per-bounce counter block has room for
  size_t budget_slots = 0;   // camera samples one wave may hold (from the last memory query), for budget_light_samples
  uint32_t budget_light_samples = 0;
  // timing of the last render
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span {
    int kernel;
    size_t e0, e1;
  };
  std::vector<Span> spans;
  float kernel_ms[K_NUM] = {0};
  uint32_t kernel_launches[K_NUM] = {0};
};

namespace {

float __int_as_float_host(int32_t v) {
  float f;
  std::memcpy(&f, &v, 4);
  return f;
}
void to3x4(const float *m16, float4 *out) {
  for (int r = 0; r < 3; ++r) out[r] = make_float4(m16[4 * r], m16[4 * r + 1], m16[4 * r + 2], m16[4 * r + 3]);
}
rpt::Box box_of_points(const float *p, size_t n) {
  rpt::Box b{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      b.mn[k] = std::fmin(b.mn[k], p[3 * i + k]);
      b.mx[k] = std::fmax(b.mx[k], p[3 * i + k]);
    }
  return b;
}
void shuffle3(const float in[3], uint32_t axis, float out[3]) {
  if (axis == RPT_AXIS_X) {
    out[0] = in[2]; out[1] = in[1]; out[2] = in[0];
  } else if (axis == RPT_AXIS_Y) {
    out[0] = in[0]; out[1] = in[2]; out[2] = in[1];
  } else {
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
  }
}
// Matrix4x4 * AABB over the 8 corners (aabb.rs:116-138)
rpt::Box transform_box(const float *m, const rpt::Box &b) {
  rpt::Box o{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
  for (int i = 0; i < 8; ++i) {
    float p[3] = {(i & 1) ? b.mx[0] : b.mn[0], (i & 2) ? b.mx[1] : b.mn[1], (i & 4) ? b.mx[2] : b.mn[2]};
    for (int r = 0; r < 3; ++r) {
      float v = m[4 * r] * p[0] + m[4 * r + 1] * p[1] + m[4 * r + 2] * p[2] + m[4 * r + 3];
      o.mn[r] = std::fmin(o.mn[r], v);
      o.mx[r] = std::fmax(o.mx[r], v);
    }
  }
  return o;
}
DevNode to_dev_node(const rpt::HostNode &h, int32_t node_offset) {
  DevNode d;
  d.lmin_lmaxx = make_float4(h.lmin[0], h.lmin[1], h.lmin[2], h.lmax[0]);
  d.lmaxyz_rminxy = make_float4(h.lmax[1], h.lmax[2], h.rmin[0], h.rmin[1]);
  d.rminz_rmax = make_float4(h.rmin[2], h.rmax[0], h.rmax[1], h.rmax[2]);
  d.children = make_int4(h.left >= 0 ? h.left + node_offset : h.left, h.right >= 0 ? h.right + node_offset : h.right, 0, 0);
  return d;
}

template <class K>
int occupancy_grid(K kernel, int threads, size_t smem, int num_sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return per_sm * num_sms;
}

// One set of wave buffers per device is kept alive across scenes: a host that creates a scene, renders and
// destroys it every frame (bench.py's e2e leg, the Rust shim) should not pay a multi-GB cudaMalloc each time.
// The caches (this one and DeviceBuffers::spare) are process-global and the ABI invites multi-threaded hosts (thread-local
// rpt_last_error, one host thread per device in rpt_multi_*): every access goes through g_cache_mu.
struct WaveCache {
  WaveBuffers wave{};
  size_t slots = 0, shadow = 0, acc = 0, counts_cap = 0;
  bool valid = false;
  float4 *film = nullptr;  // one parked film buffer (same reason: no cudaMalloc / cudaFree in a steady frame loop)
  size_t film_pixels = 0;
  // one parked stream + timing-event pool: a scene's first render on a brand-new stream with ~100 fresh events stalled the
  // host for 20-70 ms in one frame out of five (measured, tools/e2e_probe.py), with the device time unchanged
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_join = nullptr;
};
WaveCache g_wave_cache[64];

void release_wave_buffers(WaveBuffers &w) {
  void *ptrs[] = {w.paths[0], w.paths[1], w.hits, w.q_miss, w.q_diffuse, w.q_ggx, w.nee_d, w.nee_g, w.sh_a, w.sh_b, w.sh_c, w.acc, w.counts, w.work};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  w = WaveBuffers{};
}

int free_wave(RptScene *S) {
  release_wave_buffers(S->wave);
  S->wave_slots = S->wave_shadow = S->wave_acc = S->counts_cap = 0;
  return 0;
}

// Queue capacities for `valid` real entries, derived from the launch grids of the kernels that append (the
// old fixed allowance was a heuristic nothing enforced). Chunked appends lose at most 31 entries of every chunk to an
// overflow (a warp that cannot fit its <= 32 new entries pads the tail and takes a fresh chunk), and at kernel end every
// warp abandons at most one chunk per queue (per bin for the binned shadow queue). The next-path and shadow queues are
// appended to by both shade kernels (diffuse, ggx); the class lists by k_trace.
size_t path_queue_cap(const RptScene *S, size_t valid) {  // next-path queue, class lists, NEE hand-over queues
  size_t shade_warps = ((size_t)S->grid[K_SHADE_DIFFUSE] + (size_t)S->grid[K_SHADE_GGX]) * (SHADE_THREADS / 32);
  size_t trace_warps = (size_t)S->grid[K_TRACE] * (TRACE_THREADS / 32);
  size_t tail = std::max(shade_warps, trace_warps) * QCHUNK;
  // (+ min(valid, SMALL_QUEUE): launches below SMALL_QUEUE entries use 32-entry chunks, whose worst case doubles the queue)
  return valid + (valid * 31 + (QCHUNK - 31) - 1) / (QCHUNK - 31) + std::min<size_t>(valid, SMALL_QUEUE) + tail + QCHUNK;
}
size_t shadow_queue_cap(const RptScene *S, size_t valid) {
  size_t shade_warps = ((size_t)std::max(S->grid[K_SHADE_DIFFUSE], S->grid[K_NEE_DIFFUSE]) + (size_t)std::max(S->grid[K_SHADE_GGX], S->grid[K_NEE_GGX])) * (SHADE_THREADS / 32);
  const size_t bins = std::max<size_t>(NBINS, NEE_BINS), chunk = std::min<size_t>(QCHUNK_BINNED, NEE_CHUNK);
  size_t tail = shade_warps * std::max<size_t>((size_t)NBINS * QCHUNK_BINNED, (size_t)NEE_BINS * NEE_CHUNK);
  if (S->nee_mode == NEE_MODE_SORTED) tail *= 2;  // two k_nee launches per class append to the queue, each leaves its own chunk tails
  (void)bins;
  return valid + (valid * 31 + (chunk - 31) - 1) / (chunk - 31) + std::min<size_t>(valid, BIN_MIN_ITEMS) + tail + QCHUNK_BINNED;
}
// With RPT_OVERLAP=1 the wave buffers are provisioned for TWO half-waves (render_waves then runs the two halves of a wave on
// two streams): each queue holds twice the capacity of half the slots.
bool overlap_enabled() {  // RPT_OVERLAP=1 (opt-in, see render_waves)
  const char *ov = std::getenv("RPT_OVERLAP");
  return ov && ov[0] == '1';
}
size_t wave_path_cap(const RptScene *S, size_t slots) { return overlap_enabled() ? 2 * path_queue_cap(S, (slots + 1) / 2) : path_queue_cap(S, slots); }
size_t wave_shadow_cap(const RptScene *S, size_t shadow_valid) {
  return overlap_enabled() ? 2 * shadow_queue_cap(S, (shadow_valid + 1) / 2) : shadow_queue_cap(S, shadow_valid);
}
constexpr size_t kPathSlotBytes = 2 * sizeof(PathRec) + sizeof(HitRec) + 3 * sizeof(uint32_t) + 2 * sizeof(NeeRec);
constexpr size_t kShadowSlotBytes = 2 * sizeof(float4) + sizeof(uint32_t);
size_t wave_bytes(const RptScene *S, size_t slots, uint32_t light_samples) {
  return wave_path_cap(S, slots) * kPathSlotBytes + wave_shadow_cap(S, slots * light_samples) * kShadowSlotBytes + slots * sizeof(float);
}

// Called at scene destruction: park the buffers in the device's cache (keeping the larger set).
void park_wave(RptScene *S) {
  if (S->device < 0 || S->device >= 64 || !S->wave.paths[0]) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  WaveCache &c = g_wave_cache[S->device];
  if (c.valid && c.slots >= S->wave_slots && c.shadow >= S->wave_shadow && c.acc >= S->wave_acc && c.counts_cap >= S->counts_cap) {
    free_wave(S);
    return;
  }
  if (c.valid) release_wave_buffers(c.wave);
  c.wave = S->wave;
  c.slots = S->wave_slots;
  c.shadow = S->wave_shadow;
  c.acc = S->wave_acc;
  c.counts_cap = S->counts_cap;
  c.valid = true;
  S->wave = WaveBuffers{};
  S->wave_slots = S->wave_shadow = S->wave_acc = S->counts_cap = 0;
}

// true when the set parked in the device's cache would satisfy ensure_wave(S, slots, shadow_valid, bounces)
bool cached_wave_fits(const RptScene *S, size_t slots, size_t shadow_valid, size_t bounces) {
  if (S->device < 0 || S->device >= 64 || slots >= ((size_t)1 << 30)) return false;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  const WaveCache &c = g_wave_cache[S->device];
  return c.valid && wave_path_cap(S, slots) <= c.slots && wave_shadow_cap(S, shadow_valid) <= c.shadow && slots <= c.acc && bounces <= c.counts_cap;
}

// slots = camera samples in a wave; shadow_valid = NEE rays a bounce can emit; bounces = per-bounce counter rows needed
int ensure_wave(RptScene *S, size_t slots, size_t shadow_valid, size_t bounces) {
  size_t pcap = wave_path_cap(S, slots), scap = wave_shadow_cap(S, shadow_valid);
  if (pcap <= S->wave_slots && scap <= S->wave_shadow && slots <= S->wave_acc && bounces <= S->counts_cap) return 0;
  free_wave(S);
  if (S->device >= 0 && S->device < 64) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    WaveCache &c = g_wave_cache[S->device];
    if (c.valid && pcap <= c.slots && scap <= c.shadow && slots <= c.acc && bounces <= c.counts_cap) {
      S->wave = c.wave;
      S->wave_slots = c.slots;
      S->wave_shadow = c.shadow;
      S->wave_acc = c.acc;
      S->counts_cap = c.counts_cap;
      c.wave = WaveBuffers{};
      c.slots = c.shadow = c.acc = c.counts_cap = 0;
      c.valid = false;
      return 0;
    }
    if (c.valid) {  // too small for this job: release it before allocating a bigger set
      release_wave_buffers(c.wave);
      c.slots = c.shadow = c.acc = c.counts_cap = 0;
      c.valid = false;
    }
  }
  WaveBuffers &w = S->wave;
  size_t ccap = std::max<size_t>(bounces, 64);
  CUDA_TRY(cudaMalloc(&w.paths[0], pcap * sizeof(PathRec)));
  CUDA_TRY(cudaMalloc(&w.paths[1], pcap * sizeof(PathRec)));
  CUDA_TRY(cudaMalloc(&w.hits, pcap * sizeof(HitRec)));
  CUDA_TRY(cudaMalloc(&w.q_miss, pcap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&w.q_diffuse, pcap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&w.q_ggx, pcap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&w.nee_d, pcap * sizeof(NeeRec)));
  CUDA_TRY(cudaMalloc(&w.nee_g, pcap * sizeof(NeeRec)));
  CUDA_TRY(cudaMalloc(&w.sh_a, scap * sizeof(float4)));
  CUDA_TRY(cudaMalloc(&w.sh_b, scap * sizeof(float4)));
  CUDA_TRY(cudaMalloc(&w.sh_c, scap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&w.acc, slots * sizeof(float)));
  CUDA_TRY(cudaMalloc(&w.counts, 2 * (ccap + 1) * Q_COUNT * sizeof(uint32_t)));  // one block of rows per half-wave
  CUDA_TRY(cudaMalloc(&w.work, 12 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(w.work, 0, 12 * sizeof(unsigned long long)));
  S->wave_slots = pcap;
  S->wave_shadow = scap;
  S->wave_acc = slots;
  S->counts_cap = ccap;
  return 0;
}

// Per-kernel CUDA-event timing is a run-time opt-in (RPT_FLAG_KERNEL_TIMES): a plain render records two events in all.
struct Launcher {
  RptScene *S;
  bool timed;
  cudaEvent_t next_event() {
    if (S->ev_used == S->ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      S->ev_pool.push_back(e);
    }
    return S->ev_pool[S->ev_used++];
  }
  void begin(int kernel) {
    S->kernel_launches[kernel] += 1;
    if (!timed) return;
    size_t e0 = S->ev_used;
    cudaEventRecord(next_event(), S->stream);
    S->spans.push_back({kernel, e0, 0});
  }
  void end() {
    if (!timed) return;
    size_t e1 = S->ev_used;
    cudaEventRecord(next_event(), S->stream);
    S->spans.back().e1 = e1;
  }
};

int validate(const RptScene *S, const RptRenderParams *P) {
  if (!S || !P) return fail("null argument");
  if (P->width == 0 || P->height == 0) return fail("empty film");
  if (P->camera >= S->cameras.size()) return fail("camera index out of range");
  // max_bounces and light_samples are u16 in the reference (parsing/config.rs:22-23) and any value is accepted here too
  if (P->max_bounces > 65535u || P->light_samples > 65535u) return fail("max_bounces / light_samples exceed the reference's u16 range (parsing/config.rs:22-23)");
  if ((uint64_t)P->width * P->height >= (1ull << 31)) return fail("film too large");
  return 0;
}

RenderCtx make_ctx(const RptScene *S, const RptRenderParams *P) {
  RenderCtx R{};
  R.width = P->width;
  R.height = P->height;
  R.wh = P->width * P->height;
  R.min_bounces = P->min_bounces;
  R.max_bounces = P->max_bounces;
  R.light_samples = P->light_samples;
  R.only_direct = P->only_direct;
  R.lambda_lo = P->lambda_lo;
  R.lambda_hi = P->lambda_hi;
  R.seed = P->seed;
  R.cam = S->cameras[P->camera];
  {
    const char *e = std::getenv("RPT_TILED");  // RPT_TILED=0: row-major slots
    R.tiles_x = (P->width % 8u == 0u && P->height % 4u == 0u && !(e && e[0] == '0')) ? P->width / 8u : 0u;
    // floor(n * ceil(2^32 / d) / 2^32) == n / d for every n with n * d < 2^32; n < tiles_x * (height / 4)
    const uint64_t d = R.tiles_x, n_max = d * (P->height / 4u);
    R.tiles_magic = (d > 1 && n_max * d < (1ull << 32)) ? (uint32_t)(((1ull << 32) + d - 1) / d) : 0u;
  }
  return R;
}

// ---- kernel dispatch over the (traversal mode, statistics) template parameters
template <bool RAYGEN>
void launch_trace(RptScene *S, const WaveBuffers &w, cudaStream_t st, bool stats, const PathRec *in, uint32_t *cb, const RenderCtx &R, PathRec *paths_out) {
#define RPT_TRACE_LAUNCH(MODE, STATS)                                                                                                 \
  k_trace<MODE, RAYGEN, STATS><<<S->grid[K_TRACE], TRACE_THREADS, S->stack_smem, st>>>(S->dev, in, w.hits, w.q_miss, w.q_diffuse, w.q_ggx, cb, \
                                                                                       w.work, w.acc, R, paths_out)
  if (S->trav_mode == TRAV_SMALL) {
    if (stats) RPT_TRACE_LAUNCH(TRAV_SMALL, true); else RPT_TRACE_LAUNCH(TRAV_SMALL, false);
  } else if (S->trav_mode == TRAV_BVH_REFILL) {
    if (stats) RPT_TRACE_LAUNCH(TRAV_BVH_REFILL, true); else RPT_TRACE_LAUNCH(TRAV_BVH_REFILL, false);
  } else if (S->trav_mode == TRAV_BVH_FLAT) {
    if (stats) RPT_TRACE_LAUNCH(TRAV_BVH_FLAT, true); else RPT_TRACE_LAUNCH(TRAV_BVH_FLAT, false);
  } else if (S->trav_mode == TRAV_BVH4) {
    if (stats) RPT_TRACE_LAUNCH(TRAV_BVH4, true); else RPT_TRACE_LAUNCH(TRAV_BVH4, false);
  } else {
    if (stats) RPT_TRACE_LAUNCH(TRAV_BVH, true); else RPT_TRACE_LAUNCH(TRAV_BVH, false);
  }
#undef RPT_TRACE_LAUNCH
}
void launch_trace_tma(RptScene *S, const WaveBuffers &w, cudaStream_t st, bool stats, const PathRec *in, uint32_t *cb) {
  RenderCtx R{};
  if (stats)
    k_trace<TRAV_BVH_TMA, false, true><<<S->grid[K_TRACE], TRACE_THREADS, S->stack_smem, st>>>(S->dev, in, w.hits, w.q_miss, w.q_diffuse, w.q_ggx, cb, w.work, w.acc, R, nullptr);
  else
    k_trace<TRAV_BVH_TMA, false, false><<<S->grid[K_TRACE], TRACE_THREADS, S->stack_smem, st>>>(S->dev, in, w.hits, w.q_miss, w.q_diffuse, w.q_ggx, cb, w.work, w.acc, R, nullptr);
}
void launch_shadow(RptScene *S, const WaveBuffers &w, cudaStream_t st, bool stats, uint32_t *cb) {
#define RPT_SHADOW_LAUNCH(MODE, STATS) \
  k_shadow<MODE, STATS><<<S->grid[K_SHADOW], TRACE_THREADS, S->stack_smem, st>>>(S->dev, w.sh_a, w.sh_b, w.sh_c, cb, w.acc, w.work + 3)
  if (S->trav_mode == TRAV_SMALL) {
    if (stats) RPT_SHADOW_LAUNCH(TRAV_SMALL, true); else RPT_SHADOW_LAUNCH(TRAV_SMALL, false);
  } else if (S->trav_mode == TRAV_BVH_REFILL) {
    if (stats) RPT_SHADOW_LAUNCH(TRAV_BVH_REFILL, true); else RPT_SHADOW_LAUNCH(TRAV_BVH_REFILL, false);
  } else if (S->trav_mode == TRAV_BVH_FLAT) {
    if (stats) RPT_SHADOW_LAUNCH(TRAV_BVH_FLAT, true); else RPT_SHADOW_LAUNCH(TRAV_BVH_FLAT, false);
  } else if (S->trav_mode == TRAV_BVH4) {
    if (stats) RPT_SHADOW_LAUNCH(TRAV_BVH4, true); else RPT_SHADOW_LAUNCH(TRAV_BVH4, false);
  } else {
    if (stats) RPT_SHADOW_LAUNCH(TRAV_BVH, true); else RPT_SHADOW_LAUNCH(TRAV_BVH, false);
  }
#undef RPT_SHADOW_LAUNCH
}

// NEE sample generation over one class's hand-over queue, in the form the scene was given at rpt_scene_create (nee_mode).
template <uint32_t CLASS>
void launch_nee(RptScene *S, int slot, cudaStream_t st, const RenderCtx &R, uint32_t b, const NeeRec *q, uint32_t *cb, const WaveBuffers &w) {
#define RPT_NEE_LAUNCH(KIND, SORTED) k_nee<CLASS, KIND, SORTED><<<S->grid[slot], SHADE_THREADS, 0, st>>>(S->dev, R, b, q, cb, w.sh_a, w.sh_b, w.sh_c)
  switch (S->nee_mode) {
    case NEE_MODE_SORTED:
      RPT_NEE_LAUNCH(NEE_LIGHT, true);
      RPT_NEE_LAUNCH(NEE_ENV, true);
      S->kernel_launches[slot] += 1;  // (one timing span, two launches)
      break;
    case NEE_MODE_LIGHT_ONLY: RPT_NEE_LAUNCH(NEE_LIGHT, false); break;
    default: RPT_NEE_LAUNCH(NEE_BOTH, false); break;
  }
#undef RPT_NEE_LAUNCH
}

// Renders P->spp samples per pixel into S->film (un-normalised sum). Fills counters.
int render_waves(RptScene *S, const RptRenderParams *P, RptCounters *counters) {
  if (int rc = validate(S, P)) return rc;
  CUDA_TRY(cudaSetDevice(S->device));
  const size_t wh = (size_t)P->width * P->height;
  if (S->film_pixels != wh) {
    if (S->film) cudaFree(S->film);
    S->film = nullptr;
    {
      std::lock_guard<std::mutex> lk(g_cache_mu);
      WaveCache &fc = g_wave_cache[S->device & 63];
      if (fc.film && fc.film_pixels == wh) {
        S->film = fc.film;
        fc.film = nullptr;
        fc.film_pixels = 0;
      }
    }
    if (!S->film) CUDA_TRY(cudaMalloc(&S->film, wh * sizeof(float4)));
    S->film_pixels = wh;
  }
  CUDA_TRY(cudaMemsetAsync(S->film, 0, wh * sizeof(float4), S->stream));
  const bool stats = (P->flags & RPT_FLAG_BVH_STATS) != 0;
  const uint32_t max_bounces = P->only_direct ? 1u : P->max_bounces;

  // wave sizing: as many spp per wave as fit the memory budget and the 30-bit slot id. When the buffers this
  // scene already holds fit the whole job, skip the (slow) memory query: a steady-state frame loop allocates nothing.
  uint32_t spp_chunk;
  size_t want_slots = wh * (size_t)std::max<uint32_t>(P->spp, 1);
  if (want_slots < ((size_t)1 << 30) && wave_path_cap(S, want_slots) <= S->wave_slots &&
      wave_shadow_cap(S, want_slots * P->light_samples) <= S->wave_shadow && want_slots <= S->wave_acc && max_bounces <= S->counts_cap) {
    spp_chunk = std::max<uint32_t>(P->spp, 1);
  } else if (cached_wave_fits(S, want_slots, want_slots * P->light_samples, max_bounces)) {
    // a scene created for this frame (a host that uploads, renders and destroys every frame) holds nothing yet, but the set the
    // previous scene parked in the device's cache fits the whole job: ensure_wave() below takes it, no memory query needed
    spp_chunk = std::max<uint32_t>(P->spp, 1);
  } else if (S->budget_slots && S->budget_light_samples == P->light_samples && wh <= S->budget_slots) {
    // the wave budget of this scene is known from an earlier call with the same light_samples: the memory query below costs
    // tens of milliseconds in a process that holds ~100 GB (seen as 90 vs 60 ms per 128 spp furnace frame in a frame loop)
    spp_chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(P->spp ? P->spp : 1, S->budget_slots / wh));
  } else {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    size_t held = S->wave_slots * kPathSlotBytes + S->wave_shadow * kShadowSlotBytes + S->wave_acc * sizeof(float);
    if (S->device >= 0 && S->device < 64) {
      std::lock_guard<std::mutex> lk(g_cache_mu);
      const WaveCache &c = g_wave_cache[S->device];
      if (c.valid) held += c.slots * kPathSlotBytes + c.shadow * kShadowSlotBytes + c.acc * sizeof(float);
    }
    // half of what is free, at most 96 GB: a 4K x 16 spp frame (133 M paths, ~45 GB of queues) is still one wave on a 180 GB B200
    size_t budget = std::min<size_t>((size_t)((free_b + held) * 0.5), (size_t)96 << 30);
    // the queue capacities are affine in the slot count: fixed part (end-of-kernel chunk tails) + per-slot part
    size_t fixed = wave_bytes(S, 0, P->light_samples);
    size_t per_1k = wave_bytes(S, 1024, P->light_samples) - fixed;
    if (budget <= fixed) return fail("not enough device memory for the wave queues");
    size_t max_slots = std::min<size_t>((budget - fixed) / per_1k * 1024, (size_t)1 << 30);
    if (wh > max_slots) return fail("film does not fit one wave");
    S->budget_slots = max_slots;
    S->budget_light_samples = P->light_samples;
    spp_chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(P->spp ? P->spp : 1, max_slots / wh));
  }
  if (const char *e = std::getenv("RPT_WAVE_SLOTS_MAX")) {  // test hook: force several waves on a job that would fit one
    size_t cap = std::strtoull(e, nullptr, 10);
    if (cap >= wh) spp_chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(spp_chunk, cap / wh));
  }
  size_t slots = wh * spp_chunk;
  if (int rc = ensure_wave(S, slots, slots * P->light_samples, max_bounces)) {
    S->budget_slots = 0;  // (the cached budget may be stale: the next call queries the memory again)
    return rc;
  }

  S->ev_used = 0;
  S->spans.clear();
  std::memset(S->kernel_ms, 0, sizeof(S->kernel_ms));
  std::memset(S->kernel_launches, 0, sizeof(S->kernel_launches));
  Launcher T{S, (P->flags & RPT_FLAG_KERNEL_TIMES) != 0};
  const RenderCtx R0 = make_ctx(S, P);
  const size_t rows = S->counts_cap + 1;                      // counter rows of one half-wave block
  const size_t n_counts = ((size_t)max_bounces + 1) * Q_COUNT;  // ... of which this job uses the first max_bounces + 1
  std::vector<uint32_t> h_counts(2 * rows * Q_COUNT);
  RptCounters C{};
  size_t film_smem = 3 * (size_t)S->dev.num_lambda * sizeof(float);
  // Opt-in (RPT_OVERLAP=1): two half-waves on two streams. The idea: every launch of a persistent grid ends in a ragged tail
  // where SMs run dry while the last tiles finish; with the wave cut in two independent halves (disjoint sample ranges,
  // disjoint halves of every queue) on two streams, the CTAs of the other half's pending launch could move in as soon as a
  // tail frees SMs. Measured (profiles/r02_overlap.md): slower everywhere - Cornell 21.5 vs 20.6 ms, kitchen_sink 8.1 vs
  // 6.8 ms: twice the launches, and two grids that each fill the machine only take turns. The fixed cost it was after turned
  // out to be chunk padding (see QCHUNK_SMALL). Never on when per-kernel timing is on (the spans would overlap).
  const bool overlap_ok = !T.timed && overlap_enabled();
  if (overlap_ok && !S->stream2) {
    {
      std::lock_guard<std::mutex> lk(g_cache_mu);
      WaveCache &sc = g_wave_cache[S->device & 63];
      if (sc.stream2) {
        S->stream2 = sc.stream2;
        S->ev_join = sc.ev_join;
        sc.stream2 = nullptr;
        sc.ev_join = nullptr;
      }
    }
    if (!S->stream2) {
      CUDA_TRY(cudaStreamCreateWithFlags(&S->stream2, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&S->ev_join, cudaEventDisableTiming));
    }
  }
  auto part_view = [&](int k, size_t slot_off) {
    WaveBuffers v = S->wave;
    const size_t poff = (size_t)k * (S->wave_slots / 2), soff = (size_t)k * (S->wave_shadow / 2);
    v.paths[0] += poff; v.paths[1] += poff; v.hits += poff;
    v.q_miss += poff; v.q_diffuse += poff; v.q_ggx += poff; v.nee_d += poff; v.nee_g += poff;
    v.sh_a += soff; v.sh_b += soff; v.sh_c += soff;
    v.acc += slot_off;
    v.counts += (size_t)k * rows * Q_COUNT;
    v.work += (size_t)k * 6;
    return v;
  };

  if (stats) CUDA_TRY(cudaMemsetAsync(S->wave.work, 0, 12 * sizeof(unsigned long long), S->stream));
  size_t ev_first = S->ev_used;
  cudaEventRecord(T.next_event(), S->stream);
  for (uint32_t done = 0; done < P->spp; done += spp_chunk) {
    const uint32_t chunk = std::min(spp_chunk, P->spp - done);
    const int parts = (overlap_ok && chunk >= 2) ? 2 : 1;
    struct Part {
      WaveBuffers w;
      RenderCtx R;
      cudaStream_t st;
      uint32_t bounces_run;
      bool dead;
    } part[2];
    for (int k = 0; k < parts; ++k) {
      const uint32_t spp_a = parts == 2 ? (chunk + 1) / 2 : chunk;
      const uint32_t spp_k = k == 0 ? spp_a : chunk - spp_a, first = k == 0 ? 0 : spp_a;
      Part &p = part[k];
      p.w = part_view(k, wh * (size_t)first);
      p.R = R0;
      p.R.n_slots = (uint32_t)(wh * spp_k);
      p.R.sample_base = P->spp_offset + done + first;
      p.st = k == 0 ? S->stream : S->stream2;
      p.bounces_run = 0;
      p.dead = false;
      if (k == 1) {  // the second stream starts after what the first has queued so far (film clear, the previous wave's film pass)
        CUDA_TRY(cudaEventRecord(S->ev_join, S->stream));
        CUDA_TRY(cudaStreamWaitEvent(S->stream2, S->ev_join, 0));
      }
      CUDA_TRY(cudaMemsetAsync(p.w.counts, 0, n_counts * sizeof(uint32_t), p.st));
      CUDA_TRY(cudaMemsetAsync(p.w.acc, 0, (size_t)p.R.n_slots * sizeof(float), p.st));
    }
    const bool tma = S->trav_mode == TRAV_BVH_TMA;
    if (tma)  // (the other modes generate the camera vertices inside bounce 0's k_trace)
      for (int k = 0; k < parts; ++k) {
        T.begin(K_RAYGEN);
        k_raygen<<<S->grid[K_RAYGEN], 256, 0, part[k].st>>>(S->dev, part[k].R, part[k].w.paths[0], part[k].w.counts);
        T.end();
      }
    for (uint32_t b = 0; b < max_bounces; ++b) {
      for (int k = 0; k < parts; ++k) {
        Part &p = part[k];
        if (p.dead) continue;
        const WaveBuffers &w = p.w;
        const RenderCtx &R = p.R;
        cudaStream_t st = p.st;
        // Long walks (the reference accepts max_bounces up to 65535): once past 16 bounces, look at the path count every 8
        // bounces and stop launching when the wave has died out (russian roulette empties it long before).
        if (b >= 16 && (b & 7u) == 0) {
          uint32_t alive = 0;
          CUDA_TRY(cudaMemcpyAsync(&alive, w.counts + (size_t)b * Q_COUNT + N_PATHS, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
          CUDA_TRY(cudaStreamSynchronize(st));
          if (alive == 0) {
            p.dead = true;
            continue;
          }
        }
        p.bounces_run = b + 1;
        uint32_t *cb = w.counts + (size_t)b * Q_COUNT, *cn = w.counts + (size_t)(b + 1) * Q_COUNT;
        PathRec *in = w.paths[b & 1], *out = w.paths[(b + 1) & 1];
        T.begin(K_TRACE);
        if (tma)
          launch_trace_tma(S, w, st, stats, in, cb);
        else if (b == 0)
          launch_trace<true>(S, w, st, stats, nullptr, cb, R, in);
        else
          launch_trace<false>(S, w, st, stats, in, cb, R, nullptr);
        T.end();
        if (S->dev.env_kind != RPT_ENV_CONSTANT) {  // (a Constant environment's vertices are finished inside k_trace)
          T.begin(K_SHADE_MISS);
          k_shade_miss<<<S->grid[K_SHADE_MISS], 256, 0, st>>>(S->dev, in, w.q_miss, cb, w.acc);
          T.end();
        }
        if (S->fused_shade) {
          T.begin(K_SHADE_DIFFUSE);
          k_shade_surface<Q_DIFFUSE><<<S->grid[K_SHADE_DIFFUSE], SHADE_THREADS, 0, st>>>(S->dev, R, b, in, w.hits, w.q_diffuse, cb, cn, out, w.sh_a, w.sh_b, w.sh_c, w.acc);
          T.end();
          if (S->has_ggx) {  // (no GGX material in the scene: the class list stays empty, skip its 12 launches per wave)
            T.begin(K_SHADE_GGX);
            k_shade_surface<Q_GGX><<<S->grid[K_SHADE_GGX], SHADE_THREADS, 0, st>>>(S->dev, R, b, in, w.hits, w.q_ggx, cb, cn, out, w.sh_a, w.sh_b, w.sh_c, w.acc);
            T.end();
          }
        } else {
          T.begin(K_SHADE_DIFFUSE);
          k_shade_vertex<Q_DIFFUSE><<<S->grid[K_SHADE_DIFFUSE], SHADE_THREADS, 0, st>>>(S->dev, R, b, in, w.hits, w.q_diffuse, cb, cn, out, w.nee_d, w.acc);
          T.end();
          if (S->has_ggx) {
            T.begin(K_SHADE_GGX);
            k_shade_vertex<Q_GGX><<<S->grid[K_SHADE_GGX], SHADE_THREADS, 0, st>>>(S->dev, R, b, in, w.hits, w.q_ggx, cb, cn, out, w.nee_g, w.acc);
            T.end();
          }
          if (P->light_samples > 0) {
            T.begin(K_NEE_DIFFUSE);
            launch_nee<Q_DIFFUSE>(S, K_NEE_DIFFUSE, st, R, b, w.nee_d, cb, w);
            T.end();
            if (S->has_ggx) {
              T.begin(K_NEE_GGX);
              launch_nee<Q_GGX>(S, K_NEE_GGX, st, R, b, w.nee_g, cb, w);
              T.end();
            }
          }
        }
        if (P->light_samples > 0) {
          T.begin(K_SHADOW);
          launch_shadow(S, w, st, stats, cb);
          T.end();
        }
      }
    }
    if (parts == 2) {  // join: the film pass of the wave reads both halves' energies
      CUDA_TRY(cudaEventRecord(S->ev_join, S->stream2));
      CUDA_TRY(cudaStreamWaitEvent(S->stream, S->ev_join, 0));
    }
    RenderCtx Rw = R0;  // the whole wave: the halves' energies are contiguous in acc
    Rw.n_slots = (uint32_t)(wh * chunk);
    Rw.sample_base = P->spp_offset + done;
    T.begin(K_FILM);
    k_film<<<S->grid[K_FILM], 256, film_smem, S->stream>>>(S->dev, Rw, S->wave.acc, S->film);
    T.end();
    for (int k = 0; k < parts; ++k)  // (only the rows this job zeroed and used: the rest of the block is never written)
      CUDA_TRY(cudaMemcpyAsync(h_counts.data() + (size_t)k * rows * Q_COUNT, S->wave.counts + (size_t)k * rows * Q_COUNT, n_counts * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, S->stream));
    CUDA_TRY(cudaStreamSynchronize(S->stream));
    CUDA_TRY(cudaGetLastError());
    // Profile counters (profile.rs:1-8) from the queue sizes
    C.camera_rays += Rw.n_slots;
    C.bounce_rays += Rw.n_slots;  // the camera vertex (pt.rs:465, integrator/utils.rs:375)
    for (int k = 0; k < parts; ++k)
      for (uint32_t b = 0; b < part[k].bounces_run; ++b) {
        const uint32_t *c = &h_counts[((size_t)k * rows + b) * Q_COUNT];
        C.segments += c[N_PATHS];
        C.true_rays += c[N_PATHS] + c[N_SHADOW];
        C.env_hits += c[N_MISS];
        C.bounce_rays += c[N_MISS] + c[N_DIFFUSE] + c[N_GGX] - c[Q_NAN];
        C.shadow_rays += c[Q_SHADOW_REF];
        C.shadow_rays_traced += c[N_SHADOW];
        C.nee_vertices += c[N_NEE];
      }
  }
  size_t ev_last = S->ev_used;
  cudaEventRecord(T.next_event(), S->stream);
  unsigned long long h_work[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (stats) CUDA_TRY(cudaMemcpyAsync(h_work, S->wave.work, sizeof(h_work), cudaMemcpyDeviceToHost, S->stream));
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  C.walk_nodes = h_work[0] + h_work[6];
  C.walk_tris = h_work[1] + h_work[7];
  C.walk_insts = h_work[2] + h_work[8];
  C.shadow_nodes = h_work[3] + h_work[9];
  C.shadow_tris = h_work[4] + h_work[10];
  C.shadow_insts = h_work[5] + h_work[11];
  {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, S->ev_pool[ev_first], S->ev_pool[ev_last]);
    C.device_ms = ms;
  }
  for (auto &sp : S->spans) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, S->ev_pool[sp.e0], S->ev_pool[sp.e1]);
    S->kernel_ms[sp.kernel] += ms;
  }
  for (int k = 0; k < K_NUM; ++k) C.kernel_launches += S->kernel_launches[k];
  if (counters) *counters = C;
  return 0;
}

}  // namespace

// rpt_multi.cu reports its errors through the same thread-local string (internal: not part of include/rpt.h)
extern "C" __attribute__((visibility("hidden"))) void rpt_set_last_error(const char *msg) { g_error = msg ? msg : ""; }

// =================================================================================================
// C ABI
// =================================================================================================
// Builds the guide tables for the importance map the scene holds (called wherever its tables are installed). RPT_IMAP_GUIDES=0
// leaves the inversions on the plain binary search; an allocation failure does too.
static int build_imap_guides(RptScene *S) {
  DevScene &D = S->dev;
  D.imap_row_guide = D.imap_m_guide = nullptr;
  const char *e = std::getenv("RPT_IMAP_GUIDES");
  if ((e && e[0] == '0') || D.imap_rows == 0 || !D.imap_row_cdf || !D.imap_m_cdf) return 0;
  uint32_t *g_rows = nullptr, *g_m = nullptr;
  if (cudaMalloc(&g_rows, (size_t)D.imap_rows * RPT_IMAP_GUIDE * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&g_m, RPT_IMAP_GUIDE * sizeof(uint32_t)) != cudaSuccess) {
    cudaGetLastError();
    if (g_rows) cudaFree(g_rows);
    return 0;
  }
  S->bufs.ptrs.push_back(g_rows);
  S->bufs.ptrs.push_back(g_m);
  k_imap_guides<<<S->num_sms * 4, 256, 0, S->stream>>>(D.imap_rows, D.imap_cols, D.imap_row_cdf, g_rows);
  k_imap_guides<<<1, 256, 0, S->stream>>>(1u, D.imap_marginal_n, D.imap_m_cdf, g_m);
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  CUDA_TRY(cudaGetLastError());
  D.imap_row_guide = g_rows;
  D.imap_m_guide = g_m;
  return 0;
}

extern "C" {

const char *rpt_last_error(void) { return g_error.c_str(); }
uint32_t rpt_abi_version(void) { return RPT_ABI_VERSION; }

int rpt_device_count(int *count) {
  if (!count) return fail("null argument");
  CUDA_TRY(cudaGetDeviceCount(count));
  return 0;
}

int rpt_scene_destroy(RptScene *S) {
  if (!S) return 0;
  cudaSetDevice(S->device);
  park_wave(S);
  if (S->film) {
    bool parked = false;
    {
      std::lock_guard<std::mutex> lk(g_cache_mu);
      WaveCache &fc = g_wave_cache[S->device & 63];
      if (!fc.film) {
        fc.film = S->film;
        fc.film_pixels = S->film_pixels;
        parked = true;
      }
    }
    if (!parked) cudaFree(S->film);
  }
  S->bufs.release();
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    WaveCache &sc = g_wave_cache[S->device & 63];
    if (S->stream && !sc.stream) {  // (every entry point leaves the stream idle: nothing is pending on it)
      sc.stream = S->stream;
      sc.events.swap(S->ev_pool);
      S->stream = nullptr;
    }
    if (S->stream2 && !sc.stream2) {
      sc.stream2 = S->stream2;
      sc.ev_join = S->ev_join;
      S->stream2 = nullptr;
      S->ev_join = nullptr;
    }
  }
  if (S->stream2) cudaStreamDestroy(S->stream2);
  if (S->ev_join) cudaEventDestroy(S->ev_join);
  for (cudaEvent_t e : S->ev_pool) cudaEventDestroy(e);
  if (S->stream) cudaStreamDestroy(S->stream);
  delete S;
  return 0;
}

int rpt_scene_create(const RptSceneDesc *d, int device, RptScene **out) {
  if (!d || !out) return fail("null argument");
  if (d->abi_version != RPT_ABI_VERSION) return fail("ABI version mismatch (include/rpt.h RPT_ABI_VERSION)");
  if (d->num_instances == 0) return fail("scene has no instances");
  if (d->num_lambda < 2) return fail("num_lambda must be >= 2");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("no such CUDA device (there is no CPU fallback)");
  CUDA_TRY(cudaSetDevice(device));
  const bool timing = std::getenv("RPT_TIMING") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[rpt_scene_create] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_start).count());
    t_start = now;
  };
  RptScene *S = new RptScene();
  S->device = device;
  S->bufs.device = device;
  // (cudaGetDeviceProperties costs 2.5-2.8 ms per call on this driver: it was 90 % of scene creation)
  CUDA_TRY(cudaDeviceGetAttribute(&S->num_sms, cudaDevAttrMultiProcessorCount, device));
  auto bail = [&](int rc) {
    rpt_scene_destroy(S);
    return rc;
  };
  lap("device properties");
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    WaveCache &sc = g_wave_cache[device & 63];
    if (sc.stream) {
      S->stream = sc.stream;
      S->ev_pool.swap(sc.events);
      sc.stream = nullptr;
    }
  }
  if (!S->stream && cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail("cudaStreamCreate failed"));
  lap("stream create");

  // ---- geometry: per-mesh BLAS, then the TLAS over instance boxes
  std::vector<DevNode> nodes;
  std::vector<float4> tri_verts, tri_normals;
  struct MeshInfo {
    int32_t root;
    uint32_t tri_base;
    rpt::Box box;
    bool has_normals;
    uint32_t depth;
    uint32_t wide_need = 0;
  };
  // Four-wide trees (TRAV_BVH4) are opt-in: RPT_BVH4=1.
  const bool want_bvh4 = [] {
    const char *e = std::getenv("RPT_BVH4");
    return e && e[0] == '1';
  }();
  std::vector<rpt::WideBvh> blas_wide(d->num_meshes);
  std::vector<MeshInfo> minfo(d->num_meshes);
  std::vector<std::vector<DevNode>> blas_nodes(d->num_meshes);
  bool any_normals = false;
  for (uint32_t m = 0; m < d->num_meshes; ++m) any_normals |= d->meshes[m].normals != nullptr;
  uint32_t max_blas_depth = 0;
  for (uint32_t m = 0; m < d->num_meshes; ++m) {
    const RptMesh &M = d->meshes[m];
    if (M.num_faces == 0) return bail(fail("mesh without faces"));
    std::vector<rpt::Box> boxes(M.num_faces);
    for (uint32_t t = 0; t < M.num_faces; ++t) {
      float pts[9];
      for (int k = 0; k < 3; ++k) {
        uint32_t vi = M.indices[3 * t + k];
        if (vi >= M.num_vertices) return bail(fail("mesh index out of range"));
        std::memcpy(pts + 3 * k, M.vertices + 3 * (size_t)vi, 3 * sizeof(float));
      }
      boxes[t] = box_of_points(pts, 3);  // MeshTriangleRef::aabb (mesh.rs:57-64)
    }
    rpt::BuiltBvh bvh_ref = rpt::build_bvh(boxes);  // reference candidate order (tie-breaks)
#ifdef RPT_REFERENCE_TREE
    rpt::BuiltBvh &bvh = bvh_ref;
#else
    rpt::BuiltBvh bvh = rpt::build_bvh_sah(boxes);  // the tree the device walks
#endif
    minfo[m].tri_base = (uint32_t)(tri_verts.size() / 3);
    minfo[m].box = box_of_points(M.vertices, M.num_vertices);  // Mesh::new bounding box (mesh.rs:271-274)
    minfo[m].has_normals = M.normals != nullptr;
    minfo[m].root = bvh.root;  // relative; fixed up below
    minfo[m].depth = bvh.max_depth;
    max_blas_depth = std::max(max_blas_depth, bvh.max_depth);
    blas_nodes[m].reserve(bvh.nodes.size());
    for (auto &hn : bvh.nodes) blas_nodes[m].push_back(to_dev_node(hn, 0));
    if (want_bvh4) {
      blas_wide[m] = rpt::collapse_bvh4(bvh);
      minfo[m].wide_need = blas_wide[m].stack_need;
    }
    for (uint32_t t = 0; t < M.num_faces; ++t) {
      uint32_t mat = M.face_material ? M.face_material[t] : RPT_MAT_PACK(RPT_MAT_TAG_MATERIAL, 0);
      for (int k = 0; k < 3; ++k) {
        const float *p = M.vertices + 3 * (size_t)M.indices[3 * t + k];
        uint32_t wbits = k == 0 ? mat : (k == 1 ? bvh_ref.order[t] : 0u);
        float wf;
        std::memcpy(&wf, &wbits, 4);
        tri_verts.push_back(make_float4(p[0], p[1], p[2], wf));
        if (any_normals) {
          if (M.normals) {
            const float *nn = M.normals + 3 * (size_t)M.indices[3 * t + k];
            tri_normals.push_back(make_float4(nn[0], nn[1], nn[2], 0.0f));
          } else {
            tri_normals.push_back(make_float4(0, 0, 0, 0));
          }
        }
      }
    }
    S->stats.triangles += M.num_faces;
    S->stats.blas_nodes += bvh.nodes.size();
  }

  std::vector<rpt::Box> ibox(d->num_instances);
  rpt::Box world{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
  for (uint32_t i = 0; i < d->num_instances; ++i) {
    const RptInstance &I = d->instances[i];
    rpt::Box b;
    switch (I.kind) {
      case RPT_AGG_RECT: {  // rect.rs:58-66
        float half[3] = {I.size[0] / 2.0f, I.size[1] / 2.0f, 0.0001f}, v[3];
        shuffle3(half, I.axis, v);
        for (int k = 0; k < 3; ++k) {
          b.mn[k] = std::fmin(I.origin[k] - v[k], I.origin[k] + v[k]);
          b.mx[k] = std::fmax(I.origin[k] - v[k], I.origin[k] + v[k]);
        }
        break;
      }
      case RPT_AGG_SPHERE:
        for (int k = 0; k < 3; ++k) {
          b.mn[k] = I.origin[k] - I.size[0];
          b.mx[k] = I.origin[k] + I.size[0];
        }
        break;
      case RPT_AGG_DISK: {  // disk.rs:23-28 (half extent radius/2: reference quirk, kept)
        float v[3] = {I.size[0] / 2.0f, I.size[0] / 2.0f, 0.001f};
        for (int k = 0; k < 3; ++k) {
          b.mn[k] = I.origin[k] - v[k];
          b.mx[k] = I.origin[k] + v[k];
        }
        break;
      }
      case RPT_AGG_MESH:
        if (I.mesh < 0 || (uint32_t)I.mesh >= d->num_meshes) return bail(fail("instance references a missing mesh"));
        b = minfo[I.mesh].box;
        break;
      default: return bail(fail("unknown aggregate kind"));
    }
    if (I.has_transform) b = transform_box(I.forward, b);  // instance.rs:64-72
    ibox[i] = b;
    for (int k = 0; k < 3; ++k) {
      world.mn[k] = std::fmin(world.mn[k], b.mn[k]);
      world.mx[k] = std::fmax(world.mx[k], b.mx[k]);
    }
  }
  // Reference-order TLAS over the instance boxes: gives every instance its candidate order (tie-breaks).
  rpt::BuiltBvh ref_tlas = rpt::build_bvh(ibox);
  // Device TLAS: instances, except that an untransformed mesh instance contributes its triangles directly.
  std::vector<uint4> leaves;
  std::vector<rpt::Box> leaf_box;
  std::vector<char> flattened(d->num_instances, 0);
  for (uint32_t i = 0; i < d->num_instances; ++i) {
    const RptInstance &I = d->instances[i];
    if (I.kind == RPT_AGG_MESH && !I.has_transform) {
      const RptMesh &M = d->meshes[I.mesh];
      flattened[i] = 1;
      for (uint32_t t = 0; t < M.num_faces; ++t) {
        float pts[9];
        for (int k = 0; k < 3; ++k) std::memcpy(pts + 3 * k, M.vertices + 3 * (size_t)M.indices[3 * t + k], 3 * sizeof(float));
        leaves.push_back(make_uint4(i, t, minfo[I.mesh].tri_base + t, ref_tlas.order[i]));
        leaf_box.push_back(box_of_points(pts, 3));
      }
    } else {
      leaves.push_back(make_uint4(i, RPT_NONE, 0, ref_tlas.order[i]));
      leaf_box.push_back(ibox[i]);
    }
  }
#ifdef RPT_REFERENCE_TREE
  rpt::BuiltBvh tlas = rpt::build_bvh(leaf_box);
#else
  rpt::BuiltBvh tlas = rpt::build_bvh_sah(leaf_box);
#endif
  uint32_t needed_blas_depth = 0;
  for (uint32_t i = 0; i < d->num_instances; ++i)
    if (d->instances[i].kind == RPT_AGG_MESH && !flattened[i]) needed_blas_depth = std::max(needed_blas_depth, minfo[d->instances[i].mesh].depth);
  // one push per inner node on the path; shared-memory stack sized to the scene (16 / 32 / 64 / 128 entries per thread)
  uint32_t need = tlas.max_depth + needed_blas_depth + 2;
  S->flat_tlas = true;
  for (uint32_t i = 0; i < d->num_instances; ++i)
    if (d->instances[i].kind == RPT_AGG_MESH && !flattened[i]) S->flat_tlas = false;
  rpt::WideBvh tlas_wide;
  if (want_bvh4) {
    tlas_wide = rpt::collapse_bvh4(tlas);
    uint32_t blas_need = 0;
    for (uint32_t i = 0; i < d->num_instances; ++i)
      if (d->instances[i].kind == RPT_AGG_MESH && !flattened[i]) blas_need = std::max(blas_need, minfo[d->instances[i].mesh].wide_need);
    need = tlas_wide.stack_need + blas_need + 2;
    S->trav_mode = TRAV_BVH4;
  }
  S->dev.min_grab = (leaf_box.size() <= 8 && needed_blas_depth == 0) ? 4u : 1u;
  S->stack_entries = need <= 16 ? 16 : (need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : 0)));
  if (S->stack_entries == 0) return bail(fail("BVH deeper than the largest traversal stack (128 entries)"));
  S->stack_smem = (size_t)S->stack_entries * TRACE_THREADS * sizeof(int);
  // Small-scene mode: every leaf is a triangle in world space or an analytic instance, and there are few of them.
  std::vector<float4> small_tris;
  std::vector<uint4> small_leaves;
  std::vector<float> small_boxes;
  uint32_t small_ntri = 0;
  {
    // Opt-in (RPT_SMALL=1): measured on the B200 (profiles/r02_small_vs_bvh.md) the lockstep leaf walk does run 32 of 32
    // lanes, but it executes 74-85 warp instructions per Cornell ray against the BVH walk's 35 (coherent first bounce) to 85
    // (incoherent bounces), and any-hit NEE rays lose the BVH's early exit: the frame is 11 % slower. It stays selectable
    // and covered by the parity suite (test_small_scene_mode_equals_bvh).
    const char *e = std::getenv("RPT_SMALL");
    const bool allow = e && e[0] == '1';
    if (allow && needed_blas_depth == 0 && leaves.size() <= RPT_SMALL_MAX) {
      for (int pass = 0; pass < 2; ++pass)  // triangles first, then whole instances
        for (size_t l = 0; l < leaves.size(); ++l) {
          const uint4 &lf = leaves[l];
          const bool is_tri = lf.y != RPT_NONE;
          if (is_tri != (pass == 0)) continue;
          if (is_tri) {
            const float4 *tv = &tri_verts[3 * (size_t)lf.z];
            uint32_t tri_order;
            std::memcpy(&tri_order, &tv[1].w, 4);
            small_leaves.push_back(make_uint4(lf.x, lf.y, tri_order, lf.w));
            for (int kz = 0; kz < 3; ++kz)  // tri_shuffle(v, kz): kz 0 -> (y, z, x), 1 -> (z, x, y), 2 -> (x, y, z)
              for (int v = 0; v < 3; ++v) {
                const float4 &q = tv[v];
                small_tris.push_back(kz == 0 ? make_float4(q.y, q.z, q.x, 0.0f) : (kz == 1 ? make_float4(q.z, q.x, q.y, 0.0f) : make_float4(q.x, q.y, q.z, 0.0f)));
              }
            ++small_ntri;
          } else {
            small_leaves.push_back(make_uint4(lf.x, RPT_NONE, 0u, lf.w));
            small_boxes.insert(small_boxes.end(), leaf_box[l].mn, leaf_box[l].mn + 3);
            small_boxes.insert(small_boxes.end(), leaf_box[l].mx, leaf_box[l].mx + 3);
          }
        }
      S->trav_mode = TRAV_SMALL;
      S->stack_smem = std::max<size_t>(16, small_tris.size() * sizeof(float4));
    }
  }
  for (auto &hn : tlas.nodes) nodes.push_back(to_dev_node(hn, 0));
  S->stats.tlas_nodes = tlas.nodes.size();
  std::vector<int32_t> mesh_root(d->num_meshes);
  for (uint32_t m = 0; m < d->num_meshes; ++m) {
    int32_t off = (int32_t)nodes.size();
    for (DevNode n : blas_nodes[m]) {
      if (n.children.x >= 0) n.children.x += off;
      if (n.children.y >= 0) n.children.y += off;
      nodes.push_back(n);
    }
    mesh_root[m] = minfo[m].root >= 0 ? minfo[m].root + off : minfo[m].root;
  }

  std::vector<rpt::WideNode> nodes4;
  std::vector<int32_t> mesh_root4(d->num_meshes, 0);
  if (S->trav_mode == TRAV_BVH4) {
    nodes4 = tlas_wide.nodes;
    for (uint32_t m = 0; m < d->num_meshes; ++m) {
      const int32_t off = (int32_t)nodes4.size();
      for (rpt::WideNode n : blas_wide[m].nodes) {
        for (int k = 0; k < 4; ++k)
          if (n.child[k] >= 0) n.child[k] += off;
        nodes4.push_back(n);
      }
      mesh_root4[m] = blas_wide[m].root >= 0 ? blas_wide[m].root + off : blas_wide[m].root;
    }
  }
  std::vector<DevInstance> insts(d->num_instances);
  for (uint32_t i = 0; i < d->num_instances; ++i) {
    const RptInstance &I = d->instances[i];
    DevInstance &D = insts[i];
    std::memset(&D, 0, sizeof(D));
    to3x4(I.reverse, D.rev);
    to3x4(I.forward, D.fwd);
    D.origin_size0 = make_float4(I.origin[0], I.origin[1], I.origin[2], I.size[0]);
    D.size1 = I.size[1];
    D.flags = (I.kind & DI_KIND_MASK) | ((I.axis & 3u) << DI_AXIS_SHIFT) | (I.two_sided ? DI_TWO_SIDED : 0u) | (I.has_transform ? DI_HAS_TRANSFORM : 0u);
    D.material = I.material;
    D.order = ref_tlas.order[i];
    if (I.kind == RPT_AGG_MESH) {
      D.blas_root = mesh_root[I.mesh];
      D.blas_root4 = mesh_root4[I.mesh];
      D.tri_base = minfo[I.mesh].tri_base;
      D.has_normals = minfo[I.mesh].has_normals;
    }
  }
  for (uint32_t l = 0; l < d->num_lights; ++l) {
    if (d->lights[l] >= d->num_instances) return bail(fail("light references a missing instance"));
    if (d->instances[d->lights[l]].kind == RPT_AGG_MESH)
      return bail(fail("mesh lights are unimplemented in the reference (src/geometry/mesh.rs:362-386 todo!())"));
  }

  // light-material geometry for the two-phase NEE visibility query (k_shadow)
  std::vector<uint32_t> light_geom;
  bool light_geom_ok = true;
  for (uint32_t i = 0; i < d->num_instances; ++i) {
    const RptInstance &I = d->instances[i];
    bool override_light = I.material != RPT_MAT_NONE && RPT_MAT_IS_LIGHT(I.material);
    if (I.kind == RPT_AGG_MESH) {
      if (override_light) light_geom_ok = false;
      if (I.material == RPT_MAT_NONE && d->meshes[I.mesh].face_material)
        for (uint32_t t = 0; t < d->meshes[I.mesh].num_faces; ++t)
          if (RPT_MAT_IS_LIGHT(d->meshes[I.mesh].face_material[t])) light_geom_ok = false;
    } else if (override_light) {
      light_geom.push_back(i);
    }
  }
  if (!light_geom_ok || light_geom.size() > 8) light_geom.clear();
  std::vector<float> light_geom_box;
  for (uint32_t i : light_geom) {
    light_geom_box.insert(light_geom_box.end(), ibox[i].mn, ibox[i].mn + 3);
    light_geom_box.insert(light_geom_box.end(), ibox[i].mx, ibox[i].mx + 3);
  }

  lap("host BVH build + flatten");
  DevScene &D = S->dev;
  DeviceBuffers &B = S->bufs;
  int rc = 0;
  rc |= B.upload(nodes.data(), nodes.size(), &D.nodes);
  {
    const rpt::WideNode *dn4 = nullptr;
    rc |= B.upload(nodes4.data(), nodes4.size(), &dn4);
    D.nodes4 = reinterpret_cast<const float4 *>(dn4);
    D.tlas_root4 = tlas_wide.root;
  }
  rc |= B.upload(insts.data(), insts.size(), &D.instances);
  rc |= B.upload(leaves.data(), leaves.size(), &D.tlas_leaves);
  rc |= B.upload(tri_verts.data(), tri_verts.size(), &D.tri_verts);
  rc |= B.upload(tri_normals.data(), tri_normals.size(), &D.tri_normals);
  rc |= B.upload(d->lights, d->num_lights, &D.lights);
  rc |= B.upload(light_geom.data(), light_geom.size(), &D.light_geom);
  rc |= B.upload(light_geom_box.data(), light_geom_box.size(), &D.light_geom_box);
  rc |= B.upload(small_tris.data(), small_tris.size(), &D.small_tris);
  rc |= B.upload(small_leaves.data(), small_leaves.size(), &D.small_leaves);
  rc |= B.upload(small_boxes.data(), small_boxes.size(), &D.small_boxes);
  D.small_n = S->trav_mode == TRAV_SMALL ? (uint32_t)small_leaves.size() : 0u;
  D.small_ntri = small_ntri;
  D.num_light_geom = (uint32_t)light_geom.size();
  rc |= B.upload(d->materials, d->num_materials, &D.materials);
  S->has_ggx = false;
  for (uint32_t i = 0; i < d->num_materials; ++i) S->has_ggx |= d->materials[i].type == RPT_MATERIAL_GGX;
  rc |= B.upload(d->curve_lut, (size_t)d->num_curves * d->num_lambda, &D.curve_lut);
  rc |= B.upload(d->cie_lut, 3 * (size_t)d->num_lambda, &D.cie_lut);
  if (rc) return bail(rc);
  D.tlas_root = tlas.root;
  D.num_instances = d->num_instances;
  D.num_lights = d->num_lights;
  D.num_lambda = d->num_lambda;
  D.lut_lo = d->lut_lambda_lo;
  D.lut_hi = d->lut_lambda_hi;

  std::vector<DevTexture> texs(d->num_textures);
  size_t tex_bytes = 0;
  for (uint32_t t = 0; t < d->num_textures; ++t) {
    const RptTexture &T = d->textures[t];
    size_t n = (size_t)T.width * T.height * T.channels;
    if (B.upload(T.texels, n, &texs[t].texels)) return bail(1);
    tex_bytes += n * sizeof(float);
    texs[t].channels = T.channels;
    texs[t].width = T.width;
    texs[t].height = T.height;
    std::memcpy(texs[t].curves, T.curves, sizeof(T.curves));
  }
  {
    std::vector<float2> mat_fast(d->num_materials, make_float2(__int_as_float_host(-1), 0.0f));
    for (uint32_t i = 0; i < d->num_materials; ++i) {
      const RptMaterial &M = d->materials[i];
      if (M.type != RPT_MATERIAL_LAMBERTIAN || M.texstack < 0 || (uint32_t)M.texstack >= d->num_texstacks) continue;
      const RptTexStack &st = d->texstacks[M.texstack];
      if (st.count != 1) continue;
      const RptTexture &T = d->textures[d->texstack_textures[st.first]];
      if (T.channels == 1 && T.width == 1 && T.height == 1 && T.curves[0] >= 0) mat_fast[i] = make_float2(__int_as_float_host(T.curves[0]), T.texels[0]);
    }
    rc |= B.upload(mat_fast.data(), mat_fast.size(), &D.mat_fast);
  }
  rc |= B.upload(texs.data(), texs.size(), &D.textures);
  rc |= B.upload(d->texstack_textures, d->num_texstack_textures, &D.stack_tex);
  rc |= B.upload(d->texstacks, d->num_texstacks, &D.stacks);
  if (rc) return bail(rc);

  const RptEnvironment &E = d->environment;
  D.env_kind = E.kind;
  D.env_strength = E.strength;
  D.env_curve = E.curve;
  D.env_angular_diameter = E.angular_diameter;
  D.env_sun_dir = make_float3(E.sun_direction[0], E.sun_direction[1], E.sun_direction[2]);
  D.env_texstack = E.texstack;
  if (E.kind == RPT_ENV_HDR && E.texstack >= 0 && (uint32_t)E.texstack < d->num_texstacks) S->env_stack_count = d->texstacks[E.texstack].count;
  to3x4(E.rot_forward, D.env_rot_fwd);
  to3x4(E.rot_reverse, D.env_rot_rev);
  {
    // An unrotated environment (every shipped scene but one) takes the libm-free uv round trip; RPT_ENV_FAST=0 keeps the libm path.
    bool ident = true;
    for (int r = 0; r < 3; ++r) {
      const float *f = &D.env_rot_fwd[r].x, *v = &D.env_rot_rev[r].x;
      for (int c = 0; c < 3; ++c) ident = ident && f[c] == (r == c ? 1.0f : 0.0f) && v[c] == (r == c ? 1.0f : 0.0f);
    }
    const char *e = std::getenv("RPT_ENV_FAST");
    D.env_unrotated = ident && !(e && e[0] == '0');
  }
  if (E.kind == RPT_ENV_HDR && E.imap_rows) {
    size_t n = (size_t)E.imap_rows * E.imap_cols;
    D.imap_rows = E.imap_rows;
    D.imap_cols = E.imap_cols;
    D.imap_marginal_n = E.imap_marginal_n;
    D.imap_marginal_integral = E.imap_marginal_integral;
    rc |= B.upload(E.imap_row_pdf, n, &D.imap_row_pdf);
    rc |= B.upload(E.imap_row_cdf, n, &D.imap_row_cdf);
    rc |= B.upload(E.imap_marginal_pdf, E.imap_marginal_n, &D.imap_m_pdf);
    rc |= B.upload(E.imap_marginal_cdf, E.imap_marginal_n, &D.imap_m_cdf);
    if (rc) return bail(rc);
    tex_bytes += (2 * n + 2 * E.imap_marginal_n) * sizeof(float);
  }
  D.p_env = d->num_lights == 0 ? 1.0f : d->env_sampling_probability;  // world/mod.rs:77-80,170-176
  float span[3] = {world.mx[0] - world.mn[0], world.mx[1] - world.mn[1], world.mx[2] - world.mn[2]};
  D.world_radius = std::sqrt(span[0] * span[0] + span[1] * span[1] + span[2] * span[2]) / 2.0f;
  D.world_center = make_float3(world.mn[0] + span[0] / 2.0f, world.mn[1] + span[1] / 2.0f, world.mn[2] + span[2] / 2.0f);
  D.world_min = make_float3(world.mn[0], world.mn[1], world.mn[2]);
  D.world_inv_extent = make_float3(span[0] > 0 ? 1.0f / span[0] : 0.0f, span[1] > 0 ? 1.0f / span[1] : 0.0f, span[2] > 0 ? 1.0f / span[2] : 0.0f);
  S->cameras.assign(d->cameras, d->cameras + d->num_cameras);

  S->stats.instances = d->num_instances;
  S->stats.node_bytes = nodes.size() * sizeof(DevNode);
  S->stats.triangle_bytes = tri_verts.size() * sizeof(float4) + tri_normals.size() * sizeof(float4);
  S->stats.scene_bytes_total = S->stats.node_bytes + S->stats.triangle_bytes + insts.size() * sizeof(DevInstance) +
                               (size_t)d->num_curves * d->num_lambda * 4 + tex_bytes;

  lap("uploads");
  size_t film_smem = 3 * (size_t)d->num_lambda * sizeof(float);
  if (film_smem > 48 * 1024) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (cudaFuncSetAttribute(k_film, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)film_smem) != cudaSuccess)
      return bail(fail("num_lambda too large for the film kernel's shared-memory CIE tables"));
  }
  S->grid[K_RAYGEN] = occupancy_grid(k_raygen, 256, 0, S->num_sms);
  {
    const char *e = std::getenv("RPT_TMA_TILES");
    if (e && e[0] == '1' && S->trav_mode == TRAV_BVH) S->trav_mode = TRAV_BVH_TMA;
    // Lane refill is opt-in (RPT_REFILL=1). Measured (profiles/r02_refill_vs_tile.md): on the instanced 10 M-triangle scene
    // it lifts the traversal kernels from 6.5 to 9.8 lanes per instruction, but pays for it with re-read records, a lower
    // issue rate (long-scoreboard stalls 3.5 -> 4.2 per issue) and the per-step bookkeeping: 213 vs 214 ms per frame. On the
    // Cornell box (short rays) it loses 38 %, on the gem scene 12 %.
    const char *r = std::getenv("RPT_REFILL");
    if (r && r[0] == '1' && S->trav_mode == TRAV_BVH) S->trav_mode = TRAV_BVH_REFILL;
    const char *f = std::getenv("RPT_FLAT");
    if (S->trav_mode == TRAV_BVH && S->flat_tlas && !(f && f[0] == '0')) S->trav_mode = TRAV_BVH_FLAT;
  }
  if (S->stack_smem > 48 * 1024) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    const int sm = (int)S->stack_smem;
    cudaFuncSetAttribute(k_trace<TRAV_BVH, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_TMA, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_TMA, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_shadow<TRAV_BVH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_shadow<TRAV_BVH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_REFILL, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_REFILL, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_REFILL, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_REFILL, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_shadow<TRAV_BVH_REFILL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_shadow<TRAV_BVH_REFILL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace_rays<TRAV_BVH>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH4, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_shadow<TRAV_BVH4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_shadow<TRAV_BVH4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace_rays<TRAV_BVH4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_FLAT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_FLAT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_FLAT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace<TRAV_BVH_FLAT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_shadow<TRAV_BVH_FLAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_shadow<TRAV_BVH_FLAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_trace_rays<TRAV_BVH_FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
  }
  if (S->trav_mode == TRAV_SMALL) {
    S->grid[K_TRACE] = occupancy_grid(k_trace<TRAV_SMALL, false, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
    S->grid[K_SHADOW] = occupancy_grid(k_shadow<TRAV_SMALL, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
  } else if (S->trav_mode == TRAV_BVH_REFILL) {
    S->grid[K_TRACE] = occupancy_grid(k_trace<TRAV_BVH_REFILL, false, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
    S->grid[K_SHADOW] = occupancy_grid(k_shadow<TRAV_BVH_REFILL, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
  } else if (S->trav_mode == TRAV_BVH_FLAT) {
    S->grid[K_TRACE] = occupancy_grid(k_trace<TRAV_BVH_FLAT, false, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
    S->grid[K_SHADOW] = occupancy_grid(k_shadow<TRAV_BVH_FLAT, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
  } else if (S->trav_mode == TRAV_BVH4) {
    S->grid[K_TRACE] = occupancy_grid(k_trace<TRAV_BVH4, false, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
    S->grid[K_SHADOW] = occupancy_grid(k_shadow<TRAV_BVH4, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
  } else if (S->trav_mode == TRAV_BVH_TMA) {
    S->grid[K_TRACE] = occupancy_grid(k_trace<TRAV_BVH_TMA, false, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
    S->grid[K_SHADOW] = occupancy_grid(k_shadow<TRAV_BVH, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
  } else {
    S->grid[K_TRACE] = occupancy_grid(k_trace<TRAV_BVH, false, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
    S->grid[K_SHADOW] = occupancy_grid(k_shadow<TRAV_BVH, false>, TRACE_THREADS, S->stack_smem, S->num_sms);
  }
  S->grid[K_SHADE_MISS] = occupancy_grid(k_shade_miss, 256, 0, S->num_sms);
  {
    const char *e = std::getenv("RPT_FUSED_SHADE");
    S->fused_shade = e && e[0] == '1';
  }
  if (S->fused_shade) {
    S->grid[K_SHADE_DIFFUSE] = occupancy_grid(k_shade_surface<Q_DIFFUSE>, SHADE_THREADS, 0, S->num_sms);
    S->grid[K_SHADE_GGX] = occupancy_grid(k_shade_surface<Q_GGX>, SHADE_THREADS, 0, S->num_sms);
    S->grid[K_NEE_DIFFUSE] = S->grid[K_NEE_GGX] = 0;
  } else {
    S->grid[K_SHADE_DIFFUSE] = occupancy_grid(k_shade_vertex<Q_DIFFUSE>, SHADE_THREADS, 0, S->num_sms);
    S->grid[K_SHADE_GGX] = occupancy_grid(k_shade_vertex<Q_GGX>, SHADE_THREADS, 0, S->num_sms);
    {
      // Which form of k_nee the scene gets: light samples only (p_env = 0: the environment branch is not even compiled in),
      // both kinds sorted (0 < p_env < 1: one launch per kind over parked (vertex, sample) pairs; see k_nee), or the plain
      // two-kind loop (p_env = 1: the environment-only instantiation spills more than the loop that carries both branches).
      // RPT_NEE_SORT=0 selects the plain loop, RPT_NEE_SORT=1 the sorted form, whatever p_env is.
      const char *e = std::getenv("RPT_NEE_SORT");
      const float pe = S->dev.p_env;
      if (e && e[0] == '1') S->nee_mode = NEE_MODE_SORTED;
      else if (e && e[0] == '0') S->nee_mode = NEE_MODE_BOTH;
      else S->nee_mode = pe <= 0.0f ? NEE_MODE_LIGHT_ONLY : (pe >= 1.0f ? NEE_MODE_BOTH : NEE_MODE_SORTED);
    }
    switch (S->nee_mode) {
      case NEE_MODE_SORTED:  // (the two kinds share a grid size: the smaller of their occupancies)
        S->grid[K_NEE_DIFFUSE] = std::min(occupancy_grid(k_nee<Q_DIFFUSE, NEE_LIGHT, true>, SHADE_THREADS, 0, S->num_sms), occupancy_grid(k_nee<Q_DIFFUSE, NEE_ENV, true>, SHADE_THREADS, 0, S->num_sms));
        S->grid[K_NEE_GGX] = std::min(occupancy_grid(k_nee<Q_GGX, NEE_LIGHT, true>, SHADE_THREADS, 0, S->num_sms), occupancy_grid(k_nee<Q_GGX, NEE_ENV, true>, SHADE_THREADS, 0, S->num_sms));
        break;
      case NEE_MODE_LIGHT_ONLY:
        S->grid[K_NEE_DIFFUSE] = occupancy_grid(k_nee<Q_DIFFUSE, NEE_LIGHT, false>, SHADE_THREADS, 0, S->num_sms);
        S->grid[K_NEE_GGX] = occupancy_grid(k_nee<Q_GGX, NEE_LIGHT, false>, SHADE_THREADS, 0, S->num_sms);
        break;
      default:
        S->grid[K_NEE_DIFFUSE] = occupancy_grid(k_nee<Q_DIFFUSE, NEE_BOTH, false>, SHADE_THREADS, 0, S->num_sms);
        S->grid[K_NEE_GGX] = occupancy_grid(k_nee<Q_GGX, NEE_BOTH, false>, SHADE_THREADS, 0, S->num_sms);
        break;
    }
  }
  S->grid[K_FILM] = occupancy_grid(k_film, 256, film_smem, S->num_sms);
  lap("occupancy queries");
  if (build_imap_guides(S)) return bail(1);
  *out = S;
  return 0;
}

int rpt_render_pt_device(RptScene *S, const RptRenderParams *P, void **film_dev, RptCounters *counters) {
  if (!film_dev) return fail("null argument");
  if (int rc = render_waves(S, P, counters)) return rc;
  if (P->spp_total) {
    k_scale<<<S->num_sms * 4, 256, 0, S->stream>>>(S->film, S->film_pixels, 1.0f / (float)P->spp_total);
    CUDA_TRY(cudaStreamSynchronize(S->stream));
  }
  *film_dev = S->film;
  return 0;
}

int rpt_render_pt(RptScene *S, const RptRenderParams *P, float *film_xyzw, RptCounters *counters) {
  if (!film_xyzw) return fail("null argument");
  void *dev = nullptr;
  const bool timing = std::getenv("RPT_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  if (int rc = rpt_render_pt_device(S, P, &dev, counters)) return rc;
  auto t1 = std::chrono::steady_clock::now();
  CUDA_TRY(cudaMemcpyAsync(film_xyzw, dev, S->film_pixels * sizeof(float4), cudaMemcpyDeviceToHost, S->stream));
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  if (timing)
    fprintf(stderr, "[rpt_render_pt] render %8.3f ms (device %.3f)  film D2H %8.3f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
            counters ? counters->device_ms : 0.0f, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
  return 0;
}

int rpt_film_scale(RptScene *S, void *film_dev, uint64_t n_float4, float scale) {
  if (!S || !film_dev) return fail("null argument");
  CUDA_TRY(cudaSetDevice(S->device));
  k_scale<<<S->num_sms * 4, 256, 0, S->stream>>>(static_cast<float4 *>(film_dev), n_float4, scale);
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  return 0;
}

int rpt_trace_primary(RptScene *S, const RptRenderParams *P, uint32_t *inst, uint32_t *prim, float *t) {
  if (int rc = validate(S, P)) return rc;
  if (!inst || !prim || !t) return fail("null argument");
  CUDA_TRY(cudaSetDevice(S->device));
  size_t wh = (size_t)P->width * P->height;
  if (int rc = ensure_wave(S, wh, wh * P->light_samples, 1)) return rc;
  RenderCtx R = make_ctx(S, P);
  R.tiles_x = 0;  // (hit records are returned in slot order = pixel order)
  R.n_slots = (uint32_t)wh;
  R.sample_base = P->spp_offset;
  WaveBuffers &w = S->wave;
  CUDA_TRY(cudaMemsetAsync(w.counts, 0, 2 * Q_COUNT * sizeof(uint32_t), S->stream));
  if (S->trav_mode == TRAV_BVH_TMA) {
    k_raygen<<<S->grid[K_RAYGEN], 256, 0, S->stream>>>(S->dev, R, w.paths[0], w.counts);
    launch_trace_tma(S, w, S->stream, false, w.paths[0], w.counts);
  } else {
    launch_trace<true>(S, w, S->stream, false, nullptr, w.counts, R, w.paths[0]);
  }
  std::vector<HitRec> h(wh);
  CUDA_TRY(cudaMemcpyAsync(h.data(), w.hits, wh * sizeof(HitRec), cudaMemcpyDeviceToHost, S->stream));
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  CUDA_TRY(cudaGetLastError());
  for (size_t i = 0; i < wh; ++i) {
    inst[i] = h[i].inst;
    prim[i] = h[i].prim;
    t[i] = h[i].t;
  }
  return 0;
}

int rpt_trace_rays(RptScene *S, uint32_t n, const float *origins, const float *dirs, const float *tmax, uint32_t *inst, uint32_t *prim, float *t) {
  if (!S || !origins || !dirs || !tmax || !inst || !prim || !t) return fail("null argument");
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(S->device));
  float *d_o = nullptr, *d_d = nullptr, *d_t = nullptr;
  HitRec *d_h = nullptr;
  CUDA_TRY(cudaMalloc(&d_o, 3 * (size_t)n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_d, 3 * (size_t)n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_t, (size_t)n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_h, (size_t)n * sizeof(HitRec)));
  CUDA_TRY(cudaMemcpyAsync(d_o, origins, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, S->stream));
  CUDA_TRY(cudaMemcpyAsync(d_d, dirs, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, S->stream));
  CUDA_TRY(cudaMemcpyAsync(d_t, tmax, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, S->stream));
  if (S->trav_mode == TRAV_SMALL)
    k_trace_rays<TRAV_SMALL><<<S->grid[K_TRACE], TRACE_THREADS, S->stack_smem, S->stream>>>(S->dev, n, d_o, d_d, d_t, d_h);
  else if (S->trav_mode == TRAV_BVH_FLAT)
    k_trace_rays<TRAV_BVH_FLAT><<<S->grid[K_TRACE], TRACE_THREADS, S->stack_smem, S->stream>>>(S->dev, n, d_o, d_d, d_t, d_h);
  else if (S->trav_mode == TRAV_BVH4)
    k_trace_rays<TRAV_BVH4><<<S->grid[K_TRACE], TRACE_THREADS, S->stack_smem, S->stream>>>(S->dev, n, d_o, d_d, d_t, d_h);
  else
    k_trace_rays<TRAV_BVH><<<S->grid[K_TRACE], TRACE_THREADS, S->stack_smem, S->stream>>>(S->dev, n, d_o, d_d, d_t, d_h);
  std::vector<HitRec> h(n);
  CUDA_TRY(cudaMemcpyAsync(h.data(), d_h, (size_t)n * sizeof(HitRec), cudaMemcpyDeviceToHost, S->stream));
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  cudaFree(d_o);
  cudaFree(d_d);
  cudaFree(d_t);
  cudaFree(d_h);
  CUDA_TRY(cudaGetLastError());
  for (uint32_t i = 0; i < n; ++i) {
    inst[i] = h[i].inst;
    prim[i] = h[i].prim;
    t[i] = h[i].t;
  }
  return 0;
}

int rpt_output_film(RptScene *S, const float *film_xyzw, uint32_t width, uint32_t height, const RptOutputSettings *O, float *rgb_linear,
                    uint8_t *rgba8, float *l_w) {
  if (!S || !O || !rgba8) return fail("null argument");
  if (!(O->factor > 0.0f)) return fail("factor must be > 0 (renderer/mod.rs:26)");
  if (O->tonemapper > RPT_TONEMAP_REINHARD1 || O->colorspace > RPT_COLORSPACE_REC2020) return fail("unknown tonemapper / colour space");
  CUDA_TRY(cudaSetDevice(S->device));
  const uint64_t n = (uint64_t)width * height;
  if (n == 0) return fail("empty film");
  float4 *film = nullptr, *owned = nullptr;
  if (film_xyzw) {
    CUDA_TRY(cudaMalloc(&owned, n * sizeof(float4)));
    film = owned;
    CUDA_TRY(cudaMemcpyAsync(film, film_xyzw, n * sizeof(float4), cudaMemcpyHostToDevice, S->stream));
  } else {
    if (!S->film || S->film_pixels != n) return fail("no device-resident film of that size: render first or pass film_xyzw");
    film = S->film;
  }
  double *d_sums = nullptr;
  float *d_rgb = nullptr;
  uchar4 *d_rgba = nullptr;
  CUDA_TRY(cudaMalloc(&d_sums, 4 * sizeof(double)));
  CUDA_TRY(cudaMalloc(&d_rgba, n * sizeof(uchar4)));
  if (rgb_linear) CUDA_TRY(cudaMalloc(&d_rgb, 3 * n * sizeof(float)));
  CUDA_TRY(cudaMemsetAsync(d_sums, 0, 4 * sizeof(double), S->stream));
  float4 lw = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
  const bool x3 = !O->luminance_only && O->tonemapper != RPT_TONEMAP_CLAMP;
  if (O->tonemapper != RPT_TONEMAP_CLAMP) {
    float4 *d_logs = nullptr;
    if (x3) {
      CUDA_TRY(cudaMalloc(&d_logs, n * sizeof(float4)));
      k_out_log<<<S->num_sms * 4, 256, 0, S->stream>>>(film, n, d_logs);
      k_out_seqsum<<<1, 32, 0, S->stream>>>(reinterpret_cast<const float *>(d_logs), n, d_sums);
    } else {
      k_out_reduce<<<S->num_sms * 4, 256, 0, S->stream>>>(film, n, O->tonemapper == RPT_TONEMAP_REINHARD0 ? 0 : 1, d_sums);
    }
    double h[4];
    CUDA_TRY(cudaMemcpyAsync(h, d_sums, sizeof(h), cudaMemcpyDeviceToHost, S->stream));
    CUDA_TRY(cudaStreamSynchronize(S->stream));
    if (d_logs) cudaFree(d_logs);
    if (x3) {
      lw = make_float4(std::exp((float)h[0] / (float)n) / O->factor, std::exp((float)h[1] / (float)n) / O->factor,
                       std::exp((float)h[2] / (float)n) / O->factor, std::exp((float)h[3] / (float)n) / O->factor);
    } else {
      float v = (float)std::exp(h[1] / (double)n) / O->factor;
      lw = make_float4(v, v, v, v);
    }
  }
  k_out_map<<<S->num_sms * 4, 256, 0, S->stream>>>(film, n, *O, lw, d_rgb, d_rgba);
  CUDA_TRY(cudaMemcpyAsync(rgba8, d_rgba, n * sizeof(uchar4), cudaMemcpyDeviceToHost, S->stream));
  if (rgb_linear) CUDA_TRY(cudaMemcpyAsync(rgb_linear, d_rgb, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, S->stream));
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  CUDA_TRY(cudaGetLastError());
  if (l_w) {
    l_w[0] = lw.x; l_w[1] = lw.y; l_w[2] = lw.z; l_w[3] = lw.w;
  }
  cudaFree(d_sums);
  cudaFree(d_rgba);
  if (d_rgb) cudaFree(d_rgb);
  if (owned) cudaFree(owned);
  return 0;
}

int rpt_last_kernel_times(RptScene *S, RptKernelTime *out, uint32_t cap, uint32_t *n) {
  if (!S || !out || !n) return fail("null argument");
  uint32_t k = 0;
  for (int i = 0; i < K_NUM && k < cap; ++i) {
    if (S->kernel_launches[i] == 0) continue;
    out[k].name = S->fused_shade ? kKernelNamesFused[i] : kKernelNamesSplit[i];
    out[k].launches = S->kernel_launches[i];
    out[k].ms = S->kernel_ms[i];
    ++k;
  }
  *n = k;
  return 0;
}

int rpt_scene_bake_importance_map(RptScene *S, const RptImapBake *B, float *row_pdf, float *row_cdf, float *marginal_pdf, float *marginal_cdf,
                                  float *marginal_integral) {
  if (!S || !B) return fail("null argument");
  if (S->dev.env_kind != RPT_ENV_HDR) return fail("importance maps exist for HDR environments only (world/environment.rs:20-27)");
  if (B->rows == 0 || B->cols == 0 || B->num_samples == 0) return fail("empty importance map");
  if (!B->luminance || !B->basis) return fail("null curve tables");
  CUDA_TRY(cudaSetDevice(S->device));
  const uint32_t R = B->rows, Cn = B->cols, NS = B->num_samples;
  const size_t n = (size_t)R * Cn;
  const uint32_t ntex = S->env_stack_count;
  float *d_lum = nullptr, *d_basis = nullptr, *d_texel = nullptr, *d_rowsum = nullptr, *d_integral = nullptr;
  float *d_pdf = nullptr, *d_cdf = nullptr, *d_mpdf = nullptr, *d_mcdf = nullptr;
  CUDA_TRY(cudaMalloc(&d_lum, NS * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_basis, (size_t)4 * ntex * NS * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_texel, n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_rowsum, R * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_integral, sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_pdf, n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_cdf, n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_mpdf, R * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d_mcdf, R * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(d_lum, B->luminance, NS * sizeof(float), cudaMemcpyHostToDevice, S->stream));
  CUDA_TRY(cudaMemcpyAsync(d_basis, B->basis, (size_t)4 * ntex * NS * sizeof(float), cudaMemcpyHostToDevice, S->stream));
  const float step = (B->lambda_hi - B->lambda_lo) / (float)NS;
  k_imap_texel<<<S->num_sms * 8, 256, 0, S->stream>>>(S->dev, R, Cn, NS, step, d_lum, d_basis, d_texel);
  k_imap_rows<<<(R + 127) / 128, 128, 0, S->stream>>>(R, Cn, d_texel, d_pdf, d_cdf, d_rowsum);
  k_imap_marginal<<<1, 32, 0, S->stream>>>(R, d_rowsum, d_mpdf, d_mcdf, d_integral);
  float integral = 0.0f;
  CUDA_TRY(cudaMemcpyAsync(&integral, d_integral, sizeof(float), cudaMemcpyDeviceToHost, S->stream));
  if (row_pdf) CUDA_TRY(cudaMemcpyAsync(row_pdf, d_pdf, n * sizeof(float), cudaMemcpyDeviceToHost, S->stream));
  if (row_cdf) CUDA_TRY(cudaMemcpyAsync(row_cdf, d_cdf, n * sizeof(float), cudaMemcpyDeviceToHost, S->stream));
  if (marginal_pdf) CUDA_TRY(cudaMemcpyAsync(marginal_pdf, d_mpdf, R * sizeof(float), cudaMemcpyDeviceToHost, S->stream));
  if (marginal_cdf) CUDA_TRY(cudaMemcpyAsync(marginal_cdf, d_mcdf, R * sizeof(float), cudaMemcpyDeviceToHost, S->stream));
  CUDA_TRY(cudaStreamSynchronize(S->stream));
  CUDA_TRY(cudaGetLastError());
  cudaFree(d_lum);
  cudaFree(d_basis);
  cudaFree(d_texel);
  cudaFree(d_rowsum);
  cudaFree(d_integral);
  // install: the tables stay resident and replace whatever the scene was created with (old buffers live until destroy)
  S->bufs.ptrs.push_back(d_pdf);
  S->bufs.ptrs.push_back(d_cdf);
  S->bufs.ptrs.push_back(d_mpdf);
  S->bufs.ptrs.push_back(d_mcdf);
  DevScene &D = S->dev;
  D.imap_rows = R;
  D.imap_cols = Cn;
  D.imap_marginal_n = R;
  D.imap_marginal_integral = integral;
  D.imap_row_pdf = d_pdf;
  D.imap_row_cdf = d_cdf;
  D.imap_m_pdf = d_mpdf;
  D.imap_m_cdf = d_mcdf;
  if (marginal_integral) *marginal_integral = integral;
  return build_imap_guides(S);
}

int rpt_probe_bandwidth(int device, uint64_t bytes, uint32_t reps, int mode, double *gbps) {
  if (!gbps || bytes < 4096 || reps == 0 || mode < 0 || mode > 2) return fail("bad argument");
  CUDA_TRY(cudaSetDevice(device));
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  float4 *buf = nullptr;
  float *sink = nullptr;
  const uint64_t n16 = bytes / 16;
  CUDA_TRY(cudaMalloc(&buf, n16 * 16));
  CUDA_TRY(cudaMalloc(&sink, 4));
  CUDA_TRY(cudaMemset(buf, 0, n16 * 16));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int it = 0; it < 5; ++it) {  // first iteration warms the L2; best of the rest
    cudaEventRecord(e0);
    k_probe_bw<<<sms * 8, 256>>>(buf, buf, n16, reps, mode, sink);
    cudaEventRecord(e1);
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    double moved = mode == 1 ? (double)(n16 / 2) * 32.0 * reps : (mode == 2 ? (double)sms * 8 * 256 * 64.0 * reps : (double)n16 * 16.0 * reps);
    if (it > 0) best = std::max(best, moved / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sink);
  CUDA_TRY(cudaGetLastError());
  *gbps = best;
  return 0;
}

// Test hook (not in rpt.h): the uv -> direction -> uv round trip of an unrotated HDR environment through the libm path
// (which = 0) or the libm-free path (which = 1), n pairs, on `device`.
int rpt_debug_env_roundtrip(int device, uint32_t n, const float *u, const float *v, int which, float *u_out, float *v_out) {
  if (!u || !v || !u_out || !v_out) return fail("null argument");
  CUDA_TRY(cudaSetDevice(device));
  float *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 4 * (size_t)n * sizeof(float)));
  CUDA_TRY(cudaMemcpy(d, u, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d + n, v, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  k_debug_env_roundtrip<<<(n + 255) / 256, 256>>>(n, d, d + n, which, d + 2 * (size_t)n, d + 3 * (size_t)n);
  CUDA_TRY(cudaMemcpy(u_out, d + 2 * (size_t)n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(v_out, d + 3 * (size_t)n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(d);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int rpt_scene_stats(RptScene *S, RptSceneStats *out) {
  if (!S || !out) return fail("null argument");
  *out = S->stats;
  return 0;
}

}  // extern "C"
