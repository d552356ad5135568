import sys, os, time
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+"/tests")
import parity, torch, numpy as np
p = parity.pkg()
world, st, flat0 = parity.load_scene("cornell")
r = p.CudaRenderer(device=0)
pinned = torch.empty((st.height, st.width, 4), dtype=torch.float32).pin_memory()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    t0=time.perf_counter(); flat = p.ffi.FlatScene(world, st.wavelength_bounds[0], st.wavelength_bounds[1], 1024); t1=time.perf_counter()
    sc = p.ffi.Scene(r.lib, flat, 0); t2=time.perf_counter()
    cnt = sc.render_pt_into(st.params(seed=it), pinned.data_ptr()); t3=time.perf_counter()
    sc.close(); t4=time.perf_counter()
    print(f"flatten {1e3*(t1-t0):.2f} ms  scene_create {1e3*(t2-t1):.2f} ms  render+D2H {1e3*(t3-t2):.2f} ms (device {cnt.device_ms:.2f})  destroy {1e3*(t4-t3):.2f} ms")
