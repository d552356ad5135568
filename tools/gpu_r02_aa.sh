#!/bin/bash
# Round 2, GPU call AA: source-correlated profile of the traversal kernels on the 10 M-triangle instanced scene (C5).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
ncu --clock-control none --set full --import-source on -k regex:"k_trace|k_shadow" -c 4 -o /tmp/ncu/monkeys_trav -f python tools/profile_step.py instanced_monkeys 1 > gpurun_out/r02aa_ncu.log 2>&1
for id in 2 3; do
  ncu -i /tmp/ncu/monkeys_trav.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::$id 2>/dev/null | gzip -9 > gpurun_out/r02aa_monkeys_trav_cudasass_k$id.csv.gz
done
python tools/ncu_summary.py /tmp/ncu/monkeys_trav.ncu-rep gpurun_out/r02aa_monkeys_trav_ncu_kernels.csv > /dev/null 2>&1
ls -la gpurun_out/r02aa*; tail -3 gpurun_out/r02aa_ncu.log
