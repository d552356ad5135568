#!/bin/bash
# Round 2, GPU call S: source-correlated profile of the two shading kernels on Cornell (bounce 1).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
ncu --clock-control none --set full --import-source on -k regex:"k_shade_vertex|k_nee" -c 4 -o /tmp/ncu/cornell_shade -f python tools/profile_step.py cornell 1 > gpurun_out/r02s_ncu.log 2>&1
for id in 2 3; do
  ncu -i /tmp/ncu/cornell_shade.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::$id 2>/dev/null | gzip -9 > gpurun_out/r02s_cornell_shade_cudasass_k$id.csv.gz
done
python tools/ncu_summary.py /tmp/ncu/cornell_shade.ncu-rep gpurun_out/r02s_cornell_shade_ncu_kernels.csv > /dev/null 2>&1
ncu -i /tmp/ncu/cornell_shade.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/r02s_cornell_shade_raw.csv.gz
ls -la gpurun_out/r02s*; tail -3 gpurun_out/r02s_ncu.log
