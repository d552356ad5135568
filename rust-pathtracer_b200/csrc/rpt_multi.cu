// rpt_multi.cu — multi-GPU rendering behind the C ABI (include/rpt.h, rpt_multi_*).
//
// The reference is ONE process (src/bin/main.rs:59-68,170; Renderer::render, src/renderer/mod.rs:107-112), so the
// spp split + film exchange has to be callable from a single host thread of that process: a `CudaRenderer { devices }`
// in the Rust shim calls rpt_multi_create once and rpt_multi_render_pt per render setting.
//
// Work split: samples are i.i.d. (naive.rs:81-98) and the film is a plain sum (naive.rs:102), so every
// device renders the FULL frame for its share of the samples (Philox sample index = spp_offset + prefix, so N devices x k
// spp draw exactly the samples one device x N k spp would) and leaves an un-normalised XYZ sum in its own HBM. One host
// worker thread per device drives the wavefront loop of that device (rpt_render_pt_device).
//
// Film exchange, two selectable implementations of "one reduce of the XYZ film over NVLink":
//  * RPT_MULTI_PEER (default when every device can map the others' memory): ONE kernel per device, launched concurrently.
//    Device i owns pixel slice i of the frame: it reads that slice from all N films (N-1 of them through NVLink peer
//    loads), adds them, applies the 1/spp normalisation and stores the result straight into the root's film (a peer store
//    for i != 0). Reduce-scatter, gather and the normalisation pass are one fused kernel; every NVLink port carries
//    (N-1)/N of a film in and 1/N out, all ports busy at once, no staging buffer, no second launch.
//  * RPT_MULTI_NCCL: ncclReduce(sum, float, W*H*4) to the root inside ncclGroupStart/End on communicators from
//    ncclCommInitAll (NCCL is dlopen'ed: the library has no link-time dependency on it and single-GPU hosts never load it),
//    followed by the normalisation kernel on the root. Selected with RPT_MULTI_REDUCE=nccl, or automatically when peer
//    access is unavailable.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rpt.h"

extern "C" __attribute__((visibility("hidden"))) void rpt_set_last_error(const char *msg);  // rpt_kernels.cu: the thread-local error string of rpt_last_error()

namespace {

int mfail(const std::string &msg) {
  rpt_set_last_error(msg.c_str());
  return 1;
}
#define MCUDA_TRY(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return mfail(std::string(#expr) + ": " + cudaGetErrorString(_e));   \
  } while (0)

#define RPT_MULTI_MAX 16

struct FilmPtrs {
  const float4 *film[RPT_MULTI_MAX];
};

// Device `self` of n: out[i] = scale * sum_k film[k][i] for i in [begin, end). film[k] for k != self and `out` (unless self is
// the root) are peer mappings: the loads / stores travel over NVLink. 128-bit accesses, grid-stride, fully coalesced.
__global__ void __launch_bounds__(256) k_film_reduce_peer(FilmPtrs P, int n, uint64_t begin, uint64_t end, float scale, float4 *__restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
    float4 acc = P.film[0][i];
    for (int k = 1; k < n; ++k) {
      float4 v = P.film[k][i];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
    }
    acc.x *= scale;
    acc.y *= scale;
    acc.z *= scale;
    acc.w = 0.0f;
    out[i] = acc;
  }
}

// ---- NCCL through dlopen (declarations restated from nccl.h 2.x; the ABI of these entry points is stable across 2.x)
typedef void *ncclComm_t;
struct Nccl {
  void *handle = nullptr;
  int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
  static constexpr int kFloat = 7, kSum = 0;  // ncclFloat32, ncclSum
  bool load(std::string &why) {
    const char *env = std::getenv("RPT_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *nme : names) {
      if (!nme || !*nme) continue;
      handle = dlopen(nme, RTLD_NOW | RTLD_LOCAL);
      if (handle) break;
    }
    if (!handle) {
      why = std::string("NCCL not found (dlopen libnccl.so.2; set RPT_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
      return false;
    }
#define RPT_SYM(field, name)                                        \
  *(void **)(&field) = dlsym(handle, name);                         \
  if (!field) {                                                     \
    why = std::string("NCCL symbol missing: ") + name;              \
    return false;                                                   \
  }
    RPT_SYM(CommInitAll, "ncclCommInitAll");
    RPT_SYM(CommDestroy, "ncclCommDestroy");
    RPT_SYM(Reduce, "ncclReduce");
    RPT_SYM(GroupStart, "ncclGroupStart");
    RPT_SYM(GroupEnd, "ncclGroupEnd");
    RPT_SYM(GetErrorString, "ncclGetErrorString");
    RPT_SYM(GetVersion, "ncclGetVersion");
#undef RPT_SYM
    return true;
  }
};

}  // namespace

struct RptMulti {
  int n = 0;
  std::vector<int> devices;
  std::vector<RptScene *> scenes;
  std::vector<cudaStream_t> streams;          // one exchange stream per device
  std::vector<cudaEvent_t> ev0, ev1;          // exchange timing per device
  int method = RPT_MULTI_PEER;
  Nccl nccl;
  std::vector<ncclComm_t> comms;
  RptMultiTimes times{};
};

namespace {

void split_spp(uint32_t total, int n, int rank, uint32_t &count, uint32_t &offset) {  // remainder to the low ranks
  uint32_t base = total / (uint32_t)n, rem = total % (uint32_t)n;
  count = base + ((uint32_t)rank < rem ? 1u : 0u);
  offset = (uint32_t)rank * base + std::min<uint32_t>((uint32_t)rank, rem);
}

}  // namespace

extern "C" {

int rpt_multi_destroy(RptMulti *M) {
  if (!M) return 0;
  for (int i = 0; i < (int)M->comms.size(); ++i)
    if (M->comms[i]) M->nccl.CommDestroy(M->comms[i]);
  for (int i = 0; i < (int)M->scenes.size(); ++i) {
    if (!M->scenes[i]) continue;  // (a device the scene could not be created on: possibly not a valid ordinal at all)
    cudaSetDevice(M->devices[i]);
    if (i < (int)M->streams.size() && M->streams[i]) cudaStreamDestroy(M->streams[i]);
    if (i < (int)M->ev0.size() && M->ev0[i]) cudaEventDestroy(M->ev0[i]);
    if (i < (int)M->ev1.size() && M->ev1[i]) cudaEventDestroy(M->ev1[i]);
    rpt_scene_destroy(M->scenes[i]);
  }
  cudaGetLastError();  // nothing here may leave a stale error behind for the next call's cudaGetLastError() check
  delete M;
  return 0;
}

int rpt_multi_create(const RptSceneDesc *desc, const int *devices, int n, RptMulti **out) {
  if (!desc || !devices || !out) return mfail("null argument");
  if (n < 1 || n > RPT_MULTI_MAX) return mfail("device count must be 1..16");
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return mfail("a device is listed twice");
  RptMulti *M = new RptMulti();
  M->n = n;
  M->devices.assign(devices, devices + n);
  M->scenes.assign(n, nullptr);
  M->streams.assign(n, nullptr);
  M->ev0.assign(n, nullptr);
  M->ev1.assign(n, nullptr);
  // one replica of the scene per device, created concurrently (host BVH build + upload per device)
  std::vector<int> rc(n, 0);
  std::vector<std::string> err(n);
  {
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i)
      th.emplace_back([&, i]() {
        rc[i] = rpt_scene_create(desc, devices[i], &M->scenes[i]);
        if (rc[i]) err[i] = rpt_last_error();
      });
    for (auto &t : th) t.join();
  }
  for (int i = 0; i < n; ++i)
    if (rc[i]) {
      std::string e = "device " + std::to_string(devices[i]) + ": " + err[i];
      rpt_multi_destroy(M);
      return mfail(e);
    }
  for (int i = 0; i < n; ++i) {
    if (cudaSetDevice(devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&M->streams[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&M->ev0[i]) != cudaSuccess || cudaEventCreate(&M->ev1[i]) != cudaSuccess) {
      rpt_multi_destroy(M);
      return mfail("stream / event creation failed");
    }
  }
  // exchange method
  const char *env = std::getenv("RPT_MULTI_REDUCE");
  bool want_nccl = env && std::strcmp(env, "nccl") == 0;
  bool peer_ok = true;
  if (n > 1 && !want_nccl) {
    for (int i = 0; i < n && peer_ok; ++i)
      for (int j = 0; j < n && peer_ok; ++j) {
        if (i == j) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) peer_ok = false;
      }
    if (peer_ok)
      for (int i = 0; i < n; ++i) {
        cudaSetDevice(devices[i]);
        for (int j = 0; j < n; ++j) {
          if (i == j) continue;
          cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
          } else if (e != cudaSuccess) {
            peer_ok = false;
          }
        }
      }
  }
  M->method = (n == 1) ? RPT_MULTI_PEER : ((want_nccl || !peer_ok) ? RPT_MULTI_NCCL : RPT_MULTI_PEER);
  if (M->method == RPT_MULTI_NCCL && n > 1) {
    std::string why;
    if (!M->nccl.load(why)) {
      rpt_multi_destroy(M);
      return mfail(why);
    }
    M->comms.assign(n, nullptr);
    int r = M->nccl.CommInitAll(M->comms.data(), n, devices);
    if (r != 0) {
      std::string e = std::string("ncclCommInitAll: ") + M->nccl.GetErrorString(r);
      M->comms.clear();
      rpt_multi_destroy(M);
      return mfail(e);
    }
  }
  *out = M;
  return 0;
}

int rpt_multi_scene(RptMulti *M, int index, RptScene **scene) {
  if (!M || !scene || index < 0 || index >= M->n) return mfail("bad argument");
  *scene = M->scenes[index];
  return 0;
}

int rpt_multi_bake_importance_map(RptMulti *M, const RptImapBake *bake) {
  if (!M || !bake) return mfail("null argument");
  for (int i = 0; i < M->n; ++i)
    if (int rc = rpt_scene_bake_importance_map(M->scenes[i], bake, nullptr, nullptr, nullptr, nullptr, nullptr)) return rc;
  return 0;
}

int rpt_multi_render_pt(RptMulti *M, const RptRenderParams *P, float *film_xyzw, RptCounters *counters, RptMultiTimes *times) {
  if (!M || !P) return mfail("null argument");
  const int n = M->n;
  auto t_start = std::chrono::steady_clock::now();
  std::vector<RptCounters> C(n);
  std::vector<void *> film(n, nullptr);
  std::vector<int> rc(n, 0);
  std::vector<std::string> err(n);
  // ---- render: one host thread per device, each device its share of the samples, un-normalised sums stay in HBM
  auto work = [&](int i) {
    RptRenderParams p = *P;
    split_spp(P->spp, n, i, p.spp, p.spp_offset);
    p.spp_offset += P->spp_offset;
    p.spp_total = 0;  // leave the SUM: normalisation happens once, after the exchange
    rc[i] = rpt_render_pt_device(M->scenes[i], &p, &film[i], &C[i]);
    if (rc[i]) err[i] = rpt_last_error();
  };
  if (n == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i) th.emplace_back(work, i);
    for (auto &t : th) t.join();
  }
  for (int i = 0; i < n; ++i)
    if (rc[i]) return mfail("device " + std::to_string(M->devices[i]) + ": " + err[i]);
  auto t_rendered = std::chrono::steady_clock::now();

  // ---- exchange + normalisation
  const uint64_t wh = (uint64_t)P->width * P->height;
  const float scale = P->spp_total ? 1.0f / (float)P->spp_total : 1.0f;
  float4 *root_film = static_cast<float4 *>(film[0]);
  if (M->method == RPT_MULTI_PEER) {
    FilmPtrs ptrs{};
    for (int i = 0; i < n; ++i) ptrs.film[i] = static_cast<const float4 *>(film[i]);
    for (int i = 0; i < n; ++i) {
      uint64_t begin = wh * (uint64_t)i / (uint64_t)n, end = wh * (uint64_t)(i + 1) / (uint64_t)n;
      MCUDA_TRY(cudaSetDevice(M->devices[i]));
      int sms = 0;
      MCUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, M->devices[i]));
      // the slice's own film goes first so that the local read is film[0] of the kernel
      FilmPtrs local = ptrs;
      std::swap(local.film[0], local.film[i]);
      MCUDA_TRY(cudaEventRecord(M->ev0[i], M->streams[i]));
      k_film_reduce_peer<<<sms * 4, 256, 0, M->streams[i]>>>(local, n, begin, end, scale, root_film);
      MCUDA_TRY(cudaEventRecord(M->ev1[i], M->streams[i]));
    }
  } else {
    int r = M->nccl.GroupStart();
    for (int i = 0; i < n && r == 0; ++i) {
      MCUDA_TRY(cudaSetDevice(M->devices[i]));
      MCUDA_TRY(cudaEventRecord(M->ev0[i], M->streams[i]));
      r = M->nccl.Reduce(film[i], film[i], (size_t)wh * 4, Nccl::kFloat, Nccl::kSum, 0, M->comms[i], M->streams[i]);
    }
    int r2 = M->nccl.GroupEnd();
    if (r != 0 || r2 != 0) return mfail(std::string("ncclReduce: ") + M->nccl.GetErrorString(r != 0 ? r : r2));
    MCUDA_TRY(cudaSetDevice(M->devices[0]));
    FilmPtrs one{};
    one.film[0] = root_film;
    int sms = 0;
    MCUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, M->devices[0]));
    k_film_reduce_peer<<<sms * 4, 256, 0, M->streams[0]>>>(one, 1, 0, wh, scale, root_film);  // normalise (and clear the w lane)
    for (int i = 0; i < n; ++i) {
      MCUDA_TRY(cudaSetDevice(M->devices[i]));
      MCUDA_TRY(cudaEventRecord(M->ev1[i], M->streams[i]));
    }
  }
  float ex_ms = 0.0f;
  for (int i = 0; i < n; ++i) {
    MCUDA_TRY(cudaSetDevice(M->devices[i]));
    MCUDA_TRY(cudaStreamSynchronize(M->streams[i]));
    MCUDA_TRY(cudaGetLastError());
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, M->ev0[i], M->ev1[i]);
    ex_ms = std::max(ex_ms, ms);
  }
  auto t_exchanged = std::chrono::steady_clock::now();
  if (film_xyzw) {
    MCUDA_TRY(cudaSetDevice(M->devices[0]));
    MCUDA_TRY(cudaMemcpy(film_xyzw, root_film, wh * sizeof(float4), cudaMemcpyDeviceToHost));
  }
  auto t_end = std::chrono::steady_clock::now();

  // counters: sums over the devices; device_ms = the slowest device's render + the exchange
  RptCounters T{};
  double render_ms_max = 0.0;
  for (int i = 0; i < n; ++i) {
    const RptCounters &c = C[i];
    T.camera_rays += c.camera_rays; T.bounce_rays += c.bounce_rays; T.shadow_rays += c.shadow_rays; T.light_rays += c.light_rays;
    T.env_hits += c.env_hits; T.segments += c.segments; T.true_rays += c.true_rays; T.kernel_launches += c.kernel_launches;
    T.shadow_rays_traced += c.shadow_rays_traced;
    T.nee_vertices += c.nee_vertices;
    T.walk_nodes += c.walk_nodes; T.walk_tris += c.walk_tris; T.walk_insts += c.walk_insts;
    T.shadow_nodes += c.shadow_nodes; T.shadow_tris += c.shadow_tris; T.shadow_insts += c.shadow_insts;
    render_ms_max = std::max(render_ms_max, c.device_ms);
  }
  T.kernel_launches += (uint64_t)(M->method == RPT_MULTI_PEER ? n : 1);
  T.device_ms = render_ms_max + ex_ms;
  if (counters) *counters = T;
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  M->times.method = (uint32_t)M->method;
  M->times.devices = (uint32_t)n;
  M->times.render_device_ms_max = render_ms_max;
  M->times.exchange_device_ms = ex_ms;
  M->times.render_wall_ms = ms(t_start, t_rendered);
  M->times.exchange_wall_ms = ms(t_rendered, t_exchanged);
  M->times.download_wall_ms = ms(t_exchanged, t_end);
  if (times) *times = M->times;
  return 0;
}

int rpt_render_pt_multi(const RptSceneDesc *desc, const int *devices, int n, const RptRenderParams *params, float *film_xyzw, RptCounters *counters) {
  RptMulti *M = nullptr;
  if (int rc = rpt_multi_create(desc, devices, n, &M)) return rc;
  int rc = rpt_multi_render_pt(M, params, film_xyzw, counters, nullptr);
  std::string keep = rc ? rpt_last_error() : "";
  rpt_multi_destroy(M);
  if (rc) return mfail(keep);
  return 0;
}

}  // extern "C"
