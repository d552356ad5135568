#!/usr/bin/env python3
"""bench.py — headline benchmark of the PT hot path (BASELINE.json: Cornell box, 1080p, 16 spp).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one batch = rendering the whole 1920x1080 @ 16 spp frame
(33.2 M camera samples) of scenes/cornell.npz through the CUDA wavefront pipeline, film left on the
device. `value` = path segments per second (one segment = one walk ray traced + shaded, one iteration of
reference src/integrator/utils.rs:170), whole job over all N GPUs. Multi-GPU headline: every rank renders the
full frame with its own 16 spp (WEAK scaling, Philox sample offset = rank * 16), then ONE NCCL reduce of the
XYZ film to rank 0; the reduce is inside the timed region.

Also on the same JSON line:
  e2e            the same metric through the public API with host buffers (scene upload H2D + render + film D2H into
                 pinned memory, every step)
  roofline       dominant kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak
  cpu_baseline   the CPU oracle on a bounded sample of the same workload, all host threads (N = 1 only)
  strong         STRONG scaling of the same frame: 16 spp and 128 spp TOTAL split over the N GPUs, reduce time separate
  configs        one record per BASELINE config #2-#5 at its stated size, spp split over the N GPUs + one film reduce:
                 segments/s, ms, dominant kernel, its HBM (queue records) and L2 (BVH fetch) roofline fractions
  multi_inprocess  the same weak-scaling frame driven by ONE process through the C ABI's rpt_multi_* (one host thread per
                 device, fused NVLink peer reduce kernel / NCCL), measured by rank 0 while the other ranks wait
  clocks, gpu_launches

--impl reference: times the reference's CPU implementation of the path. The reference is Rust nightly with
two un-vendored git crates and cannot be built in this image, so this arm runs the C++ oracle restatement
(kind "port") with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import copy
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "path_segments_per_sec"
UNIT = "segments/s"
SCENE = "cornell"
CPU_SAMPLE = (960, 540)  # bounded CPU sample: same scene/spp/bounces, 1/4 of the pixels (~2-3 s per pass on 16 cores)

# BASELINE.json configs #2-#5 at their stated composition and size (configs[0] is the headline above).
CONFIGS = [
    {"id": "C2", "scene": "furnace", "what": "data/config_test_whitefurnace.toml, PT, 1024x1024 @ 128 spp"},
    {"id": "C3", "scene": "gem", "what": "moissanite gem (dispersive Cauchy dielectric) + gold/iron/copper/platinum/lead spheres in the Cornell rects, 1080p @ 1024 spp"},
    {"id": "C4", "scene": "hdri2", "what": "data/config_test_lighting_hdri.toml settings on hdri_test_2.toml (GGX gold / copper / dispersive glass, importance-sampled "
                                            "synthetic 4096x2048 HDR, 1000x1000 map baked on the device), 3840x2160 @ 128 spp"},
    {"id": "C5", "scene": "instanced_monkeys", "what": "2388 instances of data/meshes/monkey.obj (10.0 M triangles), 3840x2160 @ 16 spp"},
]


def load_peaks():
    """(HBM GB/s, source, L2 GB/s or None, source)."""
    hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            hbm, hbm_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    l2, l2_src = None, None
    p2 = os.path.join(ROOT, "profiles", "r02_peaks.json")
    if os.path.exists(p2):
        with open(p2) as f:
            l2 = float(json.load(f)["l2_peak_gbps"])
        l2_src = "measured (profiles/r02_peaks.json: rpt_probe_bandwidth, L2-resident streaming 128-bit reads)"
    return hbm, hbm_src, l2, l2_src


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_run(parity, spp: int, steps: int, warmup: int):
    """Times the CPU oracle on the bounded sample; returns (segments/s, seconds/step, threads, counters)."""
    world, st, flat = parity.load_scene(SCENE, CPU_SAMPLE[0], CPU_SAMPLE[1], spp)
    sc = parity.oracle_scene(flat)
    try:
        host_cores = len(os.sched_getaffinity(0))
    except AttributeError:
        host_cores = os.cpu_count() or 1
    sc.lib.rpto_set_num_threads(int(host_cores))  # all the host threads, whatever OMP_NUM_THREADS the launcher exported
    threads = int(sc.lib.rpto_num_threads())
    times, segs, cnt = [], 0, None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, cnt = sc.render_pt(st.params(seed=i))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            segs += cnt.segments
    sc.close()
    total = sum(times)
    return segs / total, total / max(1, steps), threads, cnt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import parity

    steps, warmup = args.steps, min(args.warmup, 1)
    _, st, _ = parity.load_scene(SCENE)
    value, sec_per_step, threads, cnt = cpu_oracle_run(parity, st.min_samples, steps, warmup)
    sample = f"{SCENE} {CPU_SAMPLE[0]}x{CPU_SAMPLE[1]} @ {st.min_samples} spp per step (1/4 of the 1080p frame's pixels, same bounces / light samples)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cornell_box_1080p_16spp_pt (BASELINE configs[0]); CPU arm runs the bounded sample below",
                   "max_bounces": st.max_bounces, "min_bounces": st.min_bounces, "light_samples": st.light_samples},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + "; C++/OpenMP oracle restatement of src/integrator/pt.rs (the Rust reference cannot be built here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_sec_reference_def": (cnt.camera_rays + cnt.bounce_rays + cnt.shadow_rays + cnt.light_rays) / sec_per_step if cnt else None,
    }
    print(json.dumps(line))
    return 0


def kernel_rooflines(c, ktimes, steps, hbm_peak, l2_peak):
    """Per kernel: algorithmic HBM bytes (the queue records it must move once) and, for the traversal kernels,
    the BVH bytes it fetches through L1/L2 (nodes * 64 + triangles * 48 + instances * 144), against CUDA-event time.
    `c` = counters of one instrumented step, ktimes = {name: {ms, launches}} summed over `steps` instrumented steps."""
    n_vertices = c.bounce_rays - c.camera_rays - c.env_hits  # surface vertices shaded (both classes)
    # vertex / continuing-path / shadow-record counts are not split per class by the counters: the two shade kernels share the
    # figure, so per-class fractions are upper bounds on scenes that use both classes
    shade_bytes = n_vertices * (4 + 64 + 16) + (c.segments - c.camera_rays) * 64 + c.shadow_rays_traced * 36
    # split pipeline: the vertex kernel gathers index + path + hit (84 B), writes the continuing path (64 B) and the NEE
    # hand-over record (64 B); the NEE kernel reads the hand-over record and writes the shadow records (36 B each)
    vertex_bytes = n_vertices * (4 + 64 + 16) + (c.segments - c.camera_rays) * 64 + c.nee_vertices * 64
    nee_bytes = c.nee_vertices * 64 + c.shadow_rays_traced * 36
    hbm_bytes = {
        "k_trace": c.segments * (64 + 16 + 4),        # path record in, hit record + class index out
        "k_shadow": c.shadow_rays_traced * (36 + 4),  # shadow record in, one 4-byte energy RED out
        "k_shade_surface<diffuse>": shade_bytes,
        "k_shade_surface<ggx>": shade_bytes,
        "k_shade_vertex<diffuse>": vertex_bytes,
        "k_shade_vertex<ggx>": vertex_bytes,
        "k_nee<diffuse>": nee_bytes,
        "k_nee<ggx>": nee_bytes,
        "k_shade_miss": c.env_hits * (4 + 64),
    }
    bvh_bytes = {
        "k_trace": c.walk_nodes * 64 + c.walk_tris * 48 + c.walk_insts * 144,
        "k_shadow": c.shadow_nodes * 64 + c.shadow_tris * 48 + c.shadow_insts * 144,
    }
    out = {}
    for name, kt in ktimes.items():
        if kt["launches"] == 0 or kt["ms"] <= 0:
            continue
        ms_step = kt["ms"] / steps
        r = {"ms_per_step": ms_step, "launches_per_step": kt["launches"] / steps}
        if name in hbm_bytes:
            r["hbm_gbps"] = hbm_bytes[name] / (ms_step * 1e-3) / 1e9
            r["hbm_frac"] = r["hbm_gbps"] / hbm_peak
        if name in bvh_bytes and bvh_bytes[name] > 0:
            r["bvh_fetch_gbps"] = bvh_bytes[name] / (ms_step * 1e-3) / 1e9
            if l2_peak:
                r["l2_frac"] = r["bvh_fetch_gbps"] / l2_peak
        out[name] = r
    return out, hbm_bytes, bvh_bytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default=SCENE)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--spp", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config records (BASELINE configs #2-#5)")
    ap.add_argument("--no-extras", action="store_true", help="skip strong scaling and the in-process multi-GPU leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np  # noqa: F401
    import torch

    import parity

    pkg = parity.pkg()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    W = max(args.warmup, 3)
    K = args.steps
    hbm_peak, hbm_src, l2_peak, l2_src = load_peaks()
    dev = f"cuda:{local_rank}"

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(values, op_name):
        if dist is None:
            return [float(x) for x in values]
        t = torch.tensor([float(x) for x in values], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op_name == "max" else dist.ReduceOp.SUM)
        return [float(x) for x in t]

    class Job:
        """One scene on this rank's GPU + the spp split / NCCL film reduce / normalisation around rpt_render_pt_device."""

        def __init__(self, name, width=None, height=None, spp=None):
            self.world, self.st, self.flat = parity.load_scene(name, width, height, spp)
            self.scene = parity.cuda_scene(self.flat, local_rank)  # (bakes an Unbaked importance map on the device)
            self.films = {}

        def step(self, seed, count, offset, total_spp, flags=0):
            """-> (counters, (e0, e1) torch events around reduce + normalise). Film stays on the device (rank 0 holds the mean)."""
            if world_size > 1:
                torch.cuda.current_stream().synchronize()  # the previous step's reduce still reads the film this pass clears
            st = self.st
            ptr, cnt = self.scene.render_pt_device(st.params(seed=seed, spp=count, spp_offset=offset, spp_total=0, flags=flags))
            film = self.films.get(ptr)
            if film is None:
                film = self.films[ptr] = pkg.renderer.device_tensor(ptr, (st.height, st.width, 4), local_rank)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if world_size > 1:
                dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                film.mul_(1.0 / total_spp)  # the mean over all samples (tiled.rs:396-398), on the device
            e1.record()
            return cnt, (e0, e1)

        def close(self):
            self.films.clear()
            self.scene.close()

    def timed(job, steps, warm, count, offset, total_spp, seed0):
        """`warm` untimed + `steps` timed steps bracketed by barrier + synchronize; max over ranks. -> dict."""
        for i in range(warm):
            job.step(seed0 + i, count, offset, total_spp)
        sync()
        gc.collect()
        gc.disable()
        t0 = time.perf_counter()
        dev_ms, segs, launches, pending, cnt = 0.0, 0, 0, [], None
        for i in range(steps):
            cnt, ev = job.step(seed0 + warm + i, count, offset, total_spp)
            dev_ms += cnt.device_ms
            pending.append(ev)
            segs += cnt.segments
            launches += cnt.kernel_launches + (1 if rank == 0 else 0)
        sync()
        wall = time.perf_counter() - t0
        gc.enable()
        red_ms = sum(e0.elapsed_time(e1) for e0, e1 in pending)
        dev_ms, wall_ms, red_ms = allreduce([dev_ms + red_ms, wall * 1e3, red_ms], "max")
        segs_all, launches_all = allreduce([segs, launches], "sum")
        return {"wall_ms_per_step": wall_ms / steps, "device_ms_per_step": dev_ms / steps, "reduce_ms_per_step": red_ms / steps,
                "segments_per_step": segs_all / steps, "value": segs_all / (wall_ms * 1e-3), "launches": int(launches_all), "last": cnt}

    def instrumented(job, steps, count, offset, total_spp, seed0):
        """The same steps with RPT_FLAG_KERNEL_TIMES | RPT_FLAG_BVH_STATS: per-kernel CUDA-event times and BVH work counters."""
        ktimes, prof_ms, cnt = {}, 0.0, None
        for i in range(steps):
            cnt, _ = job.step(seed0 + i, count, offset, total_spp, flags=3)
            prof_ms += cnt.device_ms
            for k in job.scene.kernel_times():
                a = ktimes.setdefault(k["name"], {"ms": 0.0, "launches": 0})
                a["ms"] += k["ms"]
                a["launches"] += k["launches"]
        sync()
        return ktimes, prof_ms / steps, cnt

    # ------------------------------------------------------------------------------------------ headline (weak scaling)
    head = Job(args.scene, args.width, args.height, args.spp)
    st = head.st
    spp = st.min_samples
    total_spp = spp * world_size  # weak scaling: per-GPU work fixed
    wh = st.width * st.height
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()  # started before the warm-up so that samples exist inside the (short) timed region
    for i in range(W):
        head.step(i, spp, rank * spp, total_spp)
    sync()
    if sampler:
        sampler.lines.clear()
    res = timed(head, K, 0, spp, rank * spp, total_spp, 1000)
    clocks = sampler.stop() if sampler else None
    value = res["value"]
    # Per-kernel CUDA-event times and BVH work counters are a run-time opt-in of the library: the K timed steps above run
    # without them; the same K steps are repeated with them for the per-kernel roofline, and what the instrumentation
    # itself costs is reported (device_ms_per_step_instrumented vs device_ms_per_step).
    ktimes, prof_ms, c = instrumented(head, K, spp, rank * spp, total_spp, 3000)

    # ---- e2e through the public API with host buffers (scene upload + render + film D2H, every step)
    renderer = pkg.CudaRenderer(device=local_rank)
    pinned = torch.empty((st.height, st.width, 4), dtype=torch.float32).pin_memory()
    e2e_steps = max(2, min(6 * K, 60))  # cheap (21 ms each) and dilutes the host-side stalls a shared box throws in now and then
    h2d = d2h = 0
    st_all = copy.copy(st)
    st_all.min_samples = total_spp

    def e2e_step(i):
        sc = renderer.make_scene(head.world, st.wavelength_bounds)  # H2D: the flattened World
        if world_size > 1:  # spp split + one NCCL reduce + normalise (CudaRenderer.render_sampled_distributed), film D2H on rank 0
            renderer.seed = 2000 + i
            film, cnt = renderer.render_sampled_distributed(sc, st_all, rank, world_size)
            if rank == 0:
                pinned.copy_(film)
            torch.cuda.synchronize()
            b = sc.stats()["scene_bytes_total"]
            sc.close()
            return cnt.segments, b
        cnt = sc.render_pt_into(st.params(seed=2000 + i, spp=spp, spp_offset=rank * spp, spp_total=spp), pinned.data_ptr())  # D2H: film
        b = sc.stats()["scene_bytes_total"]
        sc.close()
        return cnt.segments, b

    e2e_step(-1)  # one untimed warm-up: the first call pays for the per-device wave / film / scene-block caches
    gc.collect()
    gc.disable()  # as timeit does: a generation-2 collection over torch's object graph is a ~100 ms host stall in one step
    sync()
    te = time.perf_counter()
    e2e_segs, e2e_ms = 0, []
    for i in range(e2e_steps):
        ts = time.perf_counter()
        segs_i, h2d = e2e_step(i)
        e2e_ms.append((time.perf_counter() - ts) * 1e3)
        d2h = wh * 16
        e2e_segs += segs_i
    sync()
    e2e_t = time.perf_counter() - te
    gc.enable()
    e2e_t = allreduce([e2e_t], "max")[0]
    e2e_segs = allreduce([e2e_segs], "sum")[0]
    e2e_value = e2e_segs / e2e_t

    # ------------------------------------------------------------------------------------------ strong scaling (same frame)
    strong = []
    if not args.no_extras:
        for total in (16, 128):
            count, offset = pkg.renderer.split_spp(total, world_size, rank)
            r = timed(head, 3, 1, count, offset, total, 5000 + total)
            strong.append({"total_spp": total, "spp_per_gpu": count, "value": r["value"], "unit": UNIT, "ms_per_step": r["wall_ms_per_step"],
                           "device_ms_per_step": r["device_ms_per_step"], "reduce_ms_per_step": r["reduce_ms_per_step"]})
    stats_head = head.scene.stats()
    head.close()

    # ------------------------------------------------------------------------------------------ BASELINE configs #2-#5
    configs = []
    if not args.no_configs:
        for cfg in CONFIGS:
            job = Job(cfg["scene"])
            total = job.st.min_samples
            count, offset = pkg.renderer.split_spp(total, world_size, rank)
            kt, _, cc = instrumented(job, 1, count, offset, total, 7000)  # doubles as the warm-up
            r = timed(job, 2, 0, count, offset, total, 7100)
            roofs_c, _, _ = kernel_rooflines(cc, kt, 1, hbm_peak, l2_peak)
            dom = max(kt.items(), key=lambda kv: kv[1]["ms"])[0] if kt else None
            sst = job.scene.stats()
            configs.append({
                "id": cfg["id"], "scene": cfg["scene"], "what": cfg["what"], "film": f"{job.st.width}x{job.st.height}", "total_spp": total, "spp_per_gpu": count,
                "n_gpus": world_size, "scaling": "strong (spp split + one NCCL film reduce)", "value": r["value"], "unit": UNIT,
                "ms_per_step": r["wall_ms_per_step"], "device_ms_per_step": r["device_ms_per_step"], "reduce_ms_per_step": r["reduce_ms_per_step"],
                "samples_per_sec": job.st.width * job.st.height * total / (r["wall_ms_per_step"] * 1e-3),
                "dominant_kernel": dom, "dominant_kernel_roofline": roofs_c.get(dom), "kernel_rooflines_rank0": roofs_c,
                "scene_bytes": {"nodes": sst["node_bytes"], "triangles": sst["triangle_bytes"], "total": sst["scene_bytes_total"]},
                "segments_per_sample": r["segments_per_step"] / (job.st.width * job.st.height * total),
            })
            job.close()

    # ------------------------------------------------------------------------------------------ in-process multi-GPU (C ABI)
    multi = None
    if not args.no_extras:
        # The other ranks must leave their GPUs idle while rank 0 drives all of them from one process: they wait on the HOST
        # (a key in the c10d store), not in an NCCL barrier, whose kernel would spin on every waiting GPU.
        store = dist.distributed_c10d._get_default_store() if dist is not None else None
        sync()
        if rank == 0:
            try:
                world, st_m, flat = parity.load_scene(args.scene, args.width, args.height, args.spp)
                multi = {}
                for method in (["peer", "nccl"] if world_size > 1 else ["peer"]):
                    if method == "nccl":
                        os.environ["RPT_MULTI_REDUCE"] = "nccl"
                    ms = pkg.ffi.MultiScene(pkg.ffi.load_library(), flat, list(range(world_size)))
                    os.environ.pop("RPT_MULTI_REDUCE", None)
                    p = st_m.params(seed=9000, spp=spp * world_size, spp_total=spp * world_size)
                    ms.render_pt(p, film_ptr=pinned.data_ptr())  # warm-up
                    t0 = time.perf_counter()
                    segs_m, reps = 0, 3
                    for i in range(reps):
                        p.seed = 9001 + i
                        _, cm = ms.render_pt(p, film_ptr=pinned.data_ptr())
                        segs_m += cm.segments
                    dt = time.perf_counter() - t0
                    multi[method] = dict(ms.times.as_dict(), value_e2e=segs_m / dt, unit=UNIT, ms_per_step_e2e=dt * 1e3 / reps,
                                         note="rpt_multi_render_pt from ONE host thread: spp split, one worker thread per device, film exchange, "
                                              "normalisation and the film download into pinned host memory are all inside the call")
                    ms.close()
            except Exception as e:  # the headline must not depend on this leg
                multi = {"error": f"{type(e).__name__}: {e}"}
            if store is not None:
                store.set("rpt_multi_done", "1")
        elif store is not None:
            import datetime

            store.wait(["rpt_multi_done"], datetime.timedelta(seconds=900))
        sync()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: algorithmic HBM bytes / CUDA-event time (DESIGN.md "Kernels").
    roofs, hbm_bytes, bvh_bytes = kernel_rooflines(c, ktimes, K, hbm_peak, l2_peak)
    dominant = max(ktimes.items(), key=lambda kv: kv[1]["ms"])[0] if ktimes else None
    roofline = None
    if dominant in hbm_bytes:
        launches_per_step = ktimes[dominant]["launches"] / K
        per_launch_bytes = hbm_bytes[dominant] / launches_per_step
        per_launch_ms = ktimes[dominant]["ms"] / ktimes[dominant]["launches"]
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
        # measured DRAM traffic of the same kernel on the same workload: the committed ncu launch list (profiles/README.md)
        traffic, traffic_src, ncu_share = None, None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if args.scene == "cornell" and (args.width, args.height, args.spp) == (None, None, None) and os.path.exists(tpath):
            tj = json.load(open(tpath))
            if dominant in tj.get("kernels", {}):
                traffic = tj["kernels"][dominant]["dram_bytes_per_launch"]
                ncu_share = tj["kernels"][dominant]["share"]
                traffic_src = "profiles/ncu_traffic.json <- " + tj.get("source", "?") + " (dram__bytes_read.sum + dram__bytes_write.sum per launch, same command)"
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "traffic_source": traffic_src, "ncu_time_share": ncu_share, "peak_source": hbm_src, "bytes_per_launch": per_launch_bytes,
                    "ms_per_launch": per_launch_ms, "launches_per_step": launches_per_step,
                    "bvh_fetch_gbps_cache_level": bvh_bytes.get(dominant, 0) / launches_per_step / (per_launch_ms * 1e-3) / 1e9,
                    "l2_peak": l2_peak, "l2_peak_source": l2_src,
                    "note": "BVH + geometry = %d B, L1-resident: this kernel is bound by instruction issue / L1 latency under divergence, not by HBM; "
                            "see profiles/ for the ncu issue-slot and branch-efficiency counters" % (stats_head["node_bytes"] + stats_head["triangle_bytes"])}
    tot_k = max(1e-9, sum(x["ms"] for x in ktimes.values()))
    kernel_share = {k: v["ms"] / tot_k for k, v in ktimes.items()}
    split = any(k.startswith("k_shade_vertex") for k in ktimes)
    frame_hbm = sum(hbm_bytes[k] for k in (("k_trace", "k_shadow", "k_shade_vertex<diffuse>", "k_nee<diffuse>") if split else ("k_trace", "k_shadow", "k_shade_surface<diffuse>")))

    cpu = None
    if world_size == 1 and not args.no_cpu_baseline:
        v, sec, threads, _ = cpu_oracle_run(parity, spp, 4, 0)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.scene} {CPU_SAMPLE[0]}x{CPU_SAMPLE[1]} @ {spp} spp, 4 passes of {sec:.1f} s: 1/4 of the frame's pixels, same bounces / light samples; "
                         "C++/OpenMP oracle restatement (the Rust reference cannot be built in this image)"}

    ref_rays = c.camera_rays + c.bounce_rays + c.shadow_rays + c.light_rays
    step_s = res["wall_ms_per_step"] * 1e-3
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": K, "warmup": W,
        "ms_per_step": res["wall_ms_per_step"], "device_ms_per_step": res["device_ms_per_step"], "device_ms_per_step_instrumented": prof_ms,
        "reduce_ms_per_step": res["reduce_ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}_box_{st.width}x{st.height}_{spp}spp_pt (BASELINE configs[0]: data/config_test_cornell_box.toml, PT, 1080p @ 16 spp)",
                   "spp_per_gpu": spp, "total_spp": total_spp, "max_bounces": st.max_bounces, "min_bounces": st.min_bounces,
                   "light_samples": st.light_samples, "parallelism": f"spp-split x{world_size} + 1 NCCL film reduce (weak scaling: {spp} spp per GPU)",
                   "cache_note": "inputs larger than L2: one wave streams %.1f GB of queue records" % (frame_hbm / 1e9)},
        "samples_per_sec": world_size * wh * spp / step_s,
        "rays_per_sec_reference_def": world_size * ref_rays / step_s,
        "true_rays_per_sec": world_size * c.true_rays / step_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "rank0_step_ms": {"min": min(e2e_ms), "median": sorted(e2e_ms)[len(e2e_ms) // 2], "p90": sorted(e2e_ms)[(len(e2e_ms) * 9) // 10],
                                  "max": max(e2e_ms), "steps_over_1.5x_median": sum(1 for t in e2e_ms if t > 1.5 * sorted(e2e_ms)[len(e2e_ms) // 2])}},
        "gpu_launches": res["launches"],
        "clocks": clocks,
        "roofline": roofline,
        "frame_hbm_roofline": {"bytes_per_step": frame_hbm, "achieved": frame_hbm / (res["device_ms_per_step"] * 1e-3) / 1e9, "unit": "GB/s",
                               "frac": frame_hbm / (res["device_ms_per_step"] * 1e-3) / 1e9 / hbm_peak},
        "kernel_time_share": kernel_share,
        "kernel_rooflines": roofs,
        "strong": strong,
        "configs": configs,
        "multi_inprocess": multi,
        "cpu_baseline": cpu,
        "counters_last_step": c.as_dict(),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
