#!/usr/bin/env python3
"""profiles/r02_bench_n{1,2,4,8}.json (bench.py lines of one scaling session) -> profiles/r02_scaling.md."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load(n):
    return json.loads(open(os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")).read().strip().splitlines()[-1])
R = {n: load(n) for n in (1, 2, 4, 8)}
b = R[1]
L = []
L.append("# Round 2 scaling, final code, ONE 8xB200 box for every N (`tools/gpu_r02_p.sh`: `bench.py --gpus N --steps 10 --warmup 3` under torchrun, one rank per GPU)\n")
L.append("Scene: Cornell 1920x1080 (BASELINE configs[0]). Segments/s = walk rays traced + shaded per second, whole job. The box has %s host cores (CPU arm: `r02_bench_n1.json.cpu_baseline`).\n" % (b["cpu_baseline"]["cores"] if b.get("cpu_baseline") else "?"))
L.append("## Weak scaling (the headline `value`): 16 spp per GPU, total spp = 16 N, one NCCL film reduce + normalise per frame\n")
L.append("| GPUs | segments/s | x vs 1 GPU | ms/step (bracketed wall) | ms/step (CUDA events, max over ranks) | of which reduce + normalise + rank skew | e2e segments/s | e2e x |")
L.append("|---|---|---|---|---|---|---|---|")
for n in (1, 2, 4, 8):
    j = R[n]
    L.append(f"| {n} | {j['value']/1e9:.3f} G | {j['value']/b['value']:.2f} | {j['ms_per_step']:.2f} | {j['device_ms_per_step']:.2f} | {j['reduce_ms_per_step']:.3f} | {j['e2e']['value']/1e9:.3f} G | {j['e2e']['value']/b['e2e']['value']:.2f} |")
L.append("\ne2e = scene upload + spp split + render + NCCL reduce + normalise + film download on rank 0, every step (`CudaRenderer.render_sampled_distributed`).\n")
L.append("## Strong scaling: the SAME frame, total spp fixed, split over the GPUs (`strong` array of the bench line; 3 timed steps each)\n")
L.append("| GPUs | 16 spp total: ms/step | x | reduce ms | 128 spp total: ms/step | x | reduce ms |")
L.append("|---|---|---|---|---|---|---|")
x16 = x128 = 0
for n in (1, 2, 4, 8):
    s16, s128 = R[n]["strong"]
    b16, b128 = b["strong"]
    x16, x128 = b16['ms_per_step'] / s16['ms_per_step'], b128['ms_per_step'] / s128['ms_per_step']
    L.append(f"| {n} | {s16['ms_per_step']:.2f} ({s16['spp_per_gpu']} spp/GPU) | {x16:.2f} | {s16['reduce_ms_per_step']:.3f} | {s128['ms_per_step']:.2f} ({s128['spp_per_gpu']} spp/GPU) | {x128:.2f} | {s128['reduce_ms_per_step']:.3f} |")
s16_8, s128_8 = R[8]["strong"]
L.append(f"""
**What limits strong scaling.** A frame costs a fixed part plus a part proportional to the samples: on one GPU 0.69 ms for the 49 launches of
an (almost) empty frame, 2.9 ms at 1 spp, 4.25 ms at 2 spp, 20.6 ms at 16 spp, 0.54 ms per further spp (`r02_frame_time_vs_size.md`): the late
bounces of a small frame are launches over a few 100 k paths that run far below steady state. At 16 spp total on 8 GPUs each GPU renders 2 spp:
{s16_8['device_ms_per_step'] - s16_8['reduce_ms_per_step']:.1f} ms of render where {b['strong'][0]['ms_per_step']:.1f} / 8 = {b['strong'][0]['ms_per_step']/8:.1f} would be ideal, plus {s16_8['reduce_ms_per_step']:.2f} ms in which `torch.distributed.reduce` of the 33 MB film (+ `mul_`, + waiting for
the slowest rank) is not overlapped with anything: {x16:.1f}x. 128 spp total (16 spp per GPU) reaches {x128:.1f}x: a 128 spp frame on one GPU runs as waves of up to
132 M paths at {b['strong'][1]['value']/1e9:.2f} Gseg/s, a 16 spp frame at {b['strong'][0]['value']/1e9:.2f}: the same per-frame overhead, seen from the other side. Neither limiter is the
collective's bandwidth. The in-library exchange (next table) takes the reduce off the critical path almost entirely (0.06 ms).
""")
L.append("## The same weak-scaling frame driven by ONE process through the C ABI (`rpt_multi_render_pt`; `multi_inprocess` of the bench line)\n")
L.append("| GPUs | exchange | e2e segments/s (incl. film download into pinned host memory) | film exchange, device ms |")
L.append("|---|---|---|---|")
for n in (1, 2, 4, 8):
    for k, v in R[n]["multi_inprocess"].items():
        L.append(f"| {n} | {'fused NVLink peer kernel' if k == 'peer' else 'NCCL reduce + normalise'} | {v['value_e2e']/1e9:.3f} G | {v['exchange_device_ms']:.3f} |")
L.append("\n## BASELINE configs #2-#5 at their stated size, spp split over the GPUs + one film reduce (`configs` array)\n")
L.append("| config | 1 GPU ms (Gseg/s) | 2 GPUs | 4 GPUs | 8 GPUs | x at 8 | reduce ms at 8 (incl. rank skew) |")
L.append("|---|---|---|---|---|---|---|")
for i, c in enumerate(b["configs"]):
    row = [f"{c['id']} {c['scene']} {c['film']} x {c['total_spp']} spp"]
    for n in (1, 2, 4, 8):
        cc = R[n]["configs"][i]
        row.append(f"{cc['ms_per_step']:.1f} ({cc['value']/1e9:.2f})")
    row.append(f"{c['ms_per_step']/R[8]['configs'][i]['ms_per_step']:.2f}")
    row.append(f"{R[8]['configs'][i]['reduce_ms_per_step']:.2f}")
    L.append("| " + " | ".join(row) + " |")
L.append("""
C5 at 8 GPUs is 2 spp per GPU of a 4K frame: the small-frame overhead described above, plus a 133 MB film to reduce (the figure includes the wait for
the slowest rank). C3 (1024 spp = 8 x 128) and C4 are spp-rich and scale at ~7.9x.
""")
open(os.path.join(ROOT, "profiles", "r02_scaling.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L))
