#!/bin/bash
# Round 2, GPU call V: source-correlated profile of the environment-kind NEE kernel on the GGX + HDR scene (C4).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
ncu --clock-control none --set full --import-source on -k regex:"k_nee" -c 4 -o /tmp/ncu/hdri2_nee -f python tools/profile_step.py hdri2 1 > gpurun_out/r02v_ncu.log 2>&1
for id in 2 4; do
  ncu -i /tmp/ncu/hdri2_nee.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::$id 2>/dev/null | gzip -9 > gpurun_out/r02v_hdri2_nee_cudasass_k$id.csv.gz
done
python tools/ncu_summary.py /tmp/ncu/hdri2_nee.ncu-rep gpurun_out/r02v_hdri2_nee_ncu_kernels.csv > /dev/null 2>&1
ls -la gpurun_out/r02v*; tail -3 gpurun_out/r02v_ncu.log
