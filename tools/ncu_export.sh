#!/bin/bash
# ncu report -> small CSVs under gpurun_out/ (the .ncu-rep itself is 5-35 MB and stays on the box: gpurun_out/ is capped at 64 MiB)
#   tools/ncu_export.sh /tmp/ncu/name.ncu-rep name [source-kernel-id ...]
rep="$1"; name="$2"; shift 2
python tools/ncu_summary.py "$rep" "gpurun_out/${name}_kernels.csv" > /dev/null 2>&1
ncu -i "$rep" --page raw --csv 2>/dev/null | gzip -9 > "gpurun_out/${name}_raw.csv.gz"
for id in "$@"; do
  ncu -i "$rep" --page source --csv --print-source sass --kernel-id :::$id 2>/dev/null | gzip -9 > "gpurun_out/${name}_source_k${id}.csv.gz"
  ncu -i "$rep" --page source --csv --print-source cuda --kernel-id :::$id 2>/dev/null | gzip -9 > "gpurun_out/${name}_cuda_k${id}.csv.gz"
done
