#!/bin/bash
# Round 2, GPU call G: NEE origin-cell grid variants (A/B).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for scene in cornell gem instanced_monkeys hdri2 kitchen_sink; do
  timeout 600 python tools/variant_bench.py $scene librpt_b200.so librpt_var_grid3c128.so librpt_var_grid4c128.so librpt_var_grid4c64.so >> gpurun_out/r02g_variants.txt 2>> gpurun_out/r02g_variants.err
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "public_api or furnace or multi_device" > gpurun_out/r02g_pytest.log 2>&1; tail -3 gpurun_out/r02g_pytest.log
set +x
cat gpurun_out/r02g_variants.txt; tail -3 gpurun_out/r02g_variants.err
