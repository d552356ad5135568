#!/bin/bash
# Round 2, GPU call Q: source-correlated profile of the traversal kernels on Cornell (bounce 1).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
ncu --clock-control none --set full --import-source on -k regex:"k_trace|k_shadow" -c 4 -o /tmp/ncu/cornell_trav -f python tools/profile_step.py cornell 1 > gpurun_out/r02q_ncu.log 2>&1
for id in 3 4; do
  ncu -i /tmp/ncu/cornell_trav.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::$id 2>/dev/null | gzip -9 > gpurun_out/r02q_cornell_trav_cudasass_k$id.csv.gz
done
ls -la gpurun_out/r02q*; tail -3 gpurun_out/r02q_ncu.log
