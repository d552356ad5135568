#!/usr/bin/env python3
"""bench.py — headline benchmark of the PT hot path (BASELINE.json: Cornell box, 1080p, 16 spp).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one batch = rendering the whole 1920x1080 @ 16 spp frame
(33.2 M camera samples) of scenes/cornell.npz through the CUDA wavefront pipeline, film left on the
device. `value` = path segments per second (one segment = one walk ray traced + shaded, one iteration of
reference src/integrator/utils.rs:170), whole job over all N GPUs. Multi-GPU: every rank renders the full
frame with its own 16 spp (weak scaling, Philox sample offset = rank * 16), then ONE NCCL reduce of the
XYZ film to rank 0; the reduce is inside the timed region.

Also on the same JSON line: `e2e` (the same metric through the public API with host buffers: scene upload
H2D + render + film D2H into pinned memory every step), `roofline` (dominant kernel, algorithmic bytes /
CUDA-event time vs the measured HBM peak), `cpu_baseline` (the CPU oracle on a bounded sample of the same
workload, all host threads), `clocks`, `gpu_launches`.

--impl reference: times the reference's CPU implementation of the path. The reference is Rust nightly with
two un-vendored git crates and cannot be built in this image, so this arm runs the C++ oracle restatement
(kind "port") with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
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


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
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

    world, st, flat = parity.load_scene(args.scene, args.width, args.height, args.spp)
    scene = parity.cuda_scene(flat, local_rank)
    spp = st.min_samples
    total_spp = spp * world_size  # weak scaling: per-GPU work fixed
    wh = st.width * st.height

    films = {}

    def step(i: int, flags: int = 0):
        """One device-resident pass; returns (counters, (e0, e1) torch events around the reduce + normalise)."""
        if world_size > 1:
            torch.cuda.current_stream().synchronize()  # the previous step's reduce still reads the film this pass clears
        ptr, cnt = scene.render_pt_device(st.params(seed=1000 + i, spp=spp, spp_offset=rank * spp, spp_total=0, flags=flags))
        film = films.get(ptr)
        if film is None:
            film = films[ptr] = pkg.renderer.device_tensor(ptr, (st.height, st.width, 4), local_rank)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world_size > 1:
            dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            film.mul_(1.0 / total_spp)  # the mean over all samples (tiled.rs:396-398), on the device
        e1.record()
        return cnt, (e0, e1)

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()  # started before the warm-up so that samples exist inside the (short) timed region
    for i in range(W):
        step(i)
    sync()
    if sampler:
        sampler.lines.clear()
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    dev_ms, segs, launches, last_cnt = 0.0, 0, 0, None
    ktimes, pending = {}, []
    for i in range(K):
        cnt, ev = step(W + i)
        dev_ms += cnt.device_ms
        pending.append(ev)
        segs += cnt.segments
        launches += cnt.kernel_launches + (1 if rank == 0 else 0)
        last_cnt = cnt
    sync()
    wall = time.perf_counter() - t0
    gc.enable()
    clocks = sampler.stop() if sampler else None
    dev_ms += sum(e0.elapsed_time(e1) for e0, e1 in pending)
    # Per-kernel CUDA-event times and BVH work counters are a run-time opt-in of the library (RPT_FLAG_KERNEL_TIMES |
    # RPT_FLAG_BVH_STATS): the K timed steps above run without them; the same K steps are repeated with them for the
    # per-kernel roofline, and the cost of the instrumentation itself is reported (device_ms_per_step_instrumented).
    prof_ms = 0.0
    for i in range(K):
        cnt, ev = step(W + K + i, flags=3)
        prof_ms += cnt.device_ms
        last_cnt = cnt
        for k in scene.kernel_times():
            a = ktimes.setdefault(k["name"], {"ms": 0.0, "launches": 0})
            a["ms"] += k["ms"]
            a["launches"] += k["launches"]
    sync()

    # max over ranks / sums over ranks
    if dist is not None:
        t = torch.tensor([dev_ms, wall * 1e3], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(t[0]), float(t[1]) / 1e3
        s = torch.tensor([segs, launches], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        segs_all, launches_all = float(s[0]), int(s[1])
    else:
        segs_all, launches_all = float(segs), launches
    # headline: the bracketed region (barrier + synchronize on both sides), max over ranks; the CUDA-event device
    # time is reported next to it (device_ms_per_step) and is what the per-kernel roofline uses.
    value = segs_all / wall

    # ---- e2e through the public API with host buffers (scene upload + render + film D2H, every step)
    renderer = pkg.CudaRenderer(device=local_rank)
    pinned = torch.empty((st.height, st.width, 4), dtype=torch.float32).pin_memory()
    e2e_steps = max(2, min(3 * K, 30))  # cheap (24 ms each) and dilutes the host-side stalls a shared box throws in now and then
    h2d = d2h = 0

    import copy

    st_all = copy.copy(st)
    st_all.min_samples = total_spp

    def e2e_step(i):
        sc = renderer.make_scene(world, st.wavelength_bounds)  # H2D: the flattened World
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
    e2e_segs = 0
    e2e_ms = []
    for i in range(e2e_steps):
        ts = time.perf_counter()
        segs_i, h2d = e2e_step(i)
        e2e_ms.append((time.perf_counter() - ts) * 1e3)
        d2h = wh * 16
        e2e_segs += segs_i
    sync()
    e2e_t = time.perf_counter() - te
    gc.enable()
    if dist is not None:
        t = torch.tensor([e2e_t], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t[0])
        s = torch.tensor([e2e_segs], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        e2e_segs = float(s[0])
    e2e_value = e2e_segs / e2e_t

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: algorithmic HBM bytes / CUDA-event time (DESIGN.md "Kernels").
    # Algorithmic HBM bytes = the queue records a kernel must read/write once (the whole scene is a few KB
    # and stays in L1/L2, so BVH node / triangle fetches are cache traffic: reported separately as bvh_fetch).
    c = last_cnt
    stats = scene.stats()
    n_vertices = c.bounce_rays - c.camera_rays - c.env_hits  # surface vertices shaded
    hbm_bytes = {
        "k_trace": c.segments * (64 + 16 + 4),                      # path record in, hit record + class index out
        "k_shadow": c.shadow_rays_traced * (36 + 4),                # shadow record in, one 4-byte energy RED out
        "k_shade_surface<diffuse>": n_vertices * (4 + 64 + 16) + (c.segments - c.camera_rays) * 64 + c.shadow_rays_traced * 36,
    }
    bvh_bytes = {
        "k_trace": c.walk_nodes * 64 + c.walk_tris * 48 + c.walk_insts * 144,
        "k_shadow": c.shadow_nodes * 64 + c.shadow_tris * 48 + c.shadow_insts * 144,
    }
    peak, peak_src = load_peaks()
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
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "ncu_time_share": ncu_share, "peak_source": peak_src, "bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                    "launches_per_step": launches_per_step,
                    "bvh_fetch_gbps_cache_level": bvh_bytes.get(dominant, 0) / launches_per_step / (per_launch_ms * 1e-3) / 1e9,
                    "note": "BVH + geometry = %d B, L1-resident: this kernel is bound by instruction issue / L1 latency under divergence, not by HBM; "
                            "see profiles/ for the ncu issue-slot and branch-efficiency counters" % (stats["node_bytes"] + stats["triangle_bytes"])}
    kernel_share = {k: v["ms"] / max(1e-9, sum(x["ms"] for x in ktimes.values())) for k, v in ktimes.items()}

    cpu = None
    if world_size == 1 and not args.no_cpu_baseline:
        v, sec, threads, _ = cpu_oracle_run(parity, spp, 4, 0)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.scene} {CPU_SAMPLE[0]}x{CPU_SAMPLE[1]} @ {spp} spp, 4 passes of {sec:.1f} s: 1/4 of the frame's pixels, same bounces / light samples; "
                         "C++/OpenMP oracle restatement (the Rust reference cannot be built in this image)"}

    ref_rays = c.camera_rays + c.bounce_rays + c.shadow_rays + c.light_rays
    step_s = wall / K
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": K, "warmup": W,
        "ms_per_step": wall * 1e3 / K, "device_ms_per_step": dev_ms / K, "device_ms_per_step_instrumented": prof_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}_box_{st.width}x{st.height}_{spp}spp_pt (BASELINE configs[0]: data/config_test_cornell_box.toml, PT, 1080p @ 16 spp)",
                   "spp_per_gpu": spp, "total_spp": total_spp, "max_bounces": st.max_bounces, "min_bounces": st.min_bounces,
                   "light_samples": st.light_samples, "parallelism": f"spp-split x{world_size} + 1 NCCL film reduce",
                   "cache_note": "inputs larger than L2: one wave streams %.1f GB of queue records" % (wh * spp * 232 / 1e9)},
        "samples_per_sec": world_size * wh * spp / step_s,
        "rays_per_sec_reference_def": world_size * ref_rays / step_s,
        "true_rays_per_sec": world_size * c.true_rays / step_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "rank0_step_ms": {"min": min(e2e_ms), "median": sorted(e2e_ms)[len(e2e_ms) // 2], "max": max(e2e_ms)}},
        "gpu_launches": launches_all,
        "clocks": clocks,
        "roofline": roofline,
        "kernel_time_share": kernel_share,
        "cpu_baseline": cpu,
        "counters_last_step": c.as_dict(),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
