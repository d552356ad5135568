#!/usr/bin/env python3
"""Converged oracle renders for parity check (c) (BASELINE.json: "converged images at high spp agree with the reference within a
stated relative-RMSE tolerance in linear XYZ"; SURVEY §8c: >= 4096 spp on 256^2 frames, tolerance from the oracle's own
two-seed noise floor).

For every BASELINE config the CPU oracle renders the full view at 256 x 256 twice with independent seeds, 2048 spp each
(two halves of a 4096 spp estimate). Committed per scene (tests/golden/converged_<scene>.npz): the 4096 spp mean XYZ film and
`floor` = relMSE(half A, half B). A GPU render with a THIRD seed at 4096 spp is then expected at relMSE ~ floor / 2 from the
committed film (tests/test_gpu_parity.py::test_converged_against_committed_oracle). Takes ~1 h of CPU here; run once, commit.
    python tools/make_converged.py [scene ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity

SCENES = ["furnace", "hdri2", "cornell", "instanced_monkeys", "gem"]
SIZE, HALF_SPP = 256, 2048
G = os.path.join(ROOT, "tests", "golden")
for name in sys.argv[1:] or SCENES:
    world, st, flat = parity.load_scene(name, SIZE, SIZE, HALF_SPP)
    sc = parity.oracle_scene(flat)
    if os.environ.get("RPT_ORACLE_THREADS"):
        sc.lib.rpto_set_num_threads(int(os.environ["RPT_ORACLE_THREADS"]))
    t = time.time()
    a, ca = sc.render_pt(st.params(seed=101))
    b, cb = sc.render_pt(st.params(seed=202))
    sc.close()
    full = ((a.astype(np.float64) + b.astype(np.float64)) / 2).astype(np.float32)[..., :3]
    # a sample whose energy is NaN poisons its pixel (the reference paints such pixels MAUVE at tonemap time, tonemap/clamp.rs:79-81);
    # the noise floor is taken over the pixels that are finite in both halves, the poisoned fraction is recorded
    ok = np.isfinite(a[..., :3]).all(axis=2) & np.isfinite(b[..., :3]).all(axis=2)
    floor = parity.rel_mse(a[ok], b[ok])
    np.savez_compressed(os.path.join(G, f"converged_{name}.npz"), film=full, floor=np.float64(floor), half_spp=HALF_SPP, seeds=np.array([101, 202]),
                        mean_xyz=full[ok].mean(axis=0), nonfinite_pixels=int((~ok).sum()),
                        segments_per_sample=(ca.segments + cb.segments) / (2.0 * SIZE * SIZE * HALF_SPP))
    print(f"{name}: {time.time() - t:.0f} s, floor relMSE(half A, half B) = {floor:.3e}, non-finite pixels {int((~ok).sum())} of {ok.size}, mean XYZ = {full[ok].mean(axis=0)}", flush=True)
