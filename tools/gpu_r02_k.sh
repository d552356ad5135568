#!/bin/bash
# Round 2, GPU call K: same-session A/B of the chunk-size policy for small launches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
for scene in cornell kitchen_sink gem; do
  timeout 600 python tools/variant_bench.py $scene librpt_var_chunk_old.so librpt_b200.so librpt_var_chunk_1m.so librpt_var_chunk_64.so >> gpurun_out/r02k_chunks.txt 2>> gpurun_out/r02k.err
done
done
timeout 300 python - >> gpurun_out/r02k_chunks.txt 2>> gpurun_out/r02k.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
for so in ("librpt_var_chunk_old.so", "librpt_b200.so", "librpt_var_chunk_old.so", "librpt_b200.so"):
    lib = p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, so))
    for name, kw in (("cornell", {}), ("cornell", {"spp": 2}), ("cornell", {"spp": 128})):
        world, st, flat = parity.load_scene(name, **kw)
        sc = p.ffi.Scene(lib, flat, 0)
        best = 1e9
        for i in range(4):
            ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0))
            if i: best = min(best, c.device_ms)
        print(f"untimed {so:26s} {name:10s} spp {st.min_samples:4d}: {best:9.3f} ms  {c.segments / best / 1e6:6.3f} Gseg/s")
        sc.close()
PY
cat gpurun_out/r02k_chunks.txt; tail -3 gpurun_out/r02k.err
