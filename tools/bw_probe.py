#!/usr/bin/env python3
"""Measured memory-system peaks of the box (rpt_probe_bandwidth): L2-resident streaming reads, HBM streaming reads, copy,
and random 64-byte gathers at BVH-sized and HBM-sized footprints. -> JSON on stdout (committed as profiles/rNN_peaks.json)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
p = parity.pkg()
lib = p.ffi.load_library()
MB = 1 << 20
out = {"how": "rpt_probe_bandwidth (csrc/rpt_kernels.cu k_probe_bw), best of 4 launches after one warm-up, CUDA events; 8 CTAs x 256 threads per SM"}
out["l2_stream_read_gbps"] = {f"{mb}MB": p.ffi.probe_bandwidth(lib, 0, mb * MB, max(1, 2048 // mb), 0) for mb in (8, 16, 32, 64, 96)}
out["hbm_stream_read_gbps"] = {"4096MB": p.ffi.probe_bandwidth(lib, 0, 4096 * MB, 1, 0)}
out["hbm_copy_gbps"] = {"4096MB": p.ffi.probe_bandwidth(lib, 0, 4096 * MB, 1, 1)}
out["gather64_gbps"] = {f"{mb}MB": p.ffi.probe_bandwidth(lib, 0, int(mb * MB), 256, 2) for mb in (0.125, 0.75, 8, 32, 64, 1024)}  # 256 gathers per thread
out["l2_peak_gbps"] = max(out["l2_stream_read_gbps"].values())
print(json.dumps(out, indent=1))
