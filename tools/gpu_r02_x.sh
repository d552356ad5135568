#!/bin/bash
# Round 2, GPU call X: which of the two device round trips agrees with the host libm (the oracle's arithmetic).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/env_roundtrip_probe.py 2000000 > gpurun_out/r02x_env_roundtrip.txt 2> gpurun_out/r02x.err
cat gpurun_out/r02x_env_roundtrip.txt; tail -3 gpurun_out/r02x.err
timeout 300 python tools/gpu_env_fast_film_diff.py > gpurun_out/r02x_films.txt 2>&1; cat gpurun_out/r02x_films.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02x_tests_all.log 2>&1
tail -4 gpurun_out/r02x_tests_all.log
