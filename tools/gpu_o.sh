#!/bin/bash
mkdir -p gpurun_out/sanitizer
S=/usr/local/cuda/bin/compute-sanitizer
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_catalog_api.py tests/test_gpu_listing.py -m gpu -q -x 2>&1 | tail -3
for mode in tile queue; do
  timeout 600 $S --tool racecheck --print-limit 20 python tools/sanitize_target.py $mode > gpurun_out/sanitizer/racecheck_${mode}.log 2>&1
  echo "racecheck $mode $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer/racecheck_${mode}.log | tail -1)"
done
timeout 600 $S --tool racecheck --print-limit 20 python tools/sanitize_target.py auto > gpurun_out/sanitizer/racecheck_auto.log 2>&1
echo "racecheck auto $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer/racecheck_auto.log | tail -1)"
timeout 600 $S --tool initcheck --print-limit 20 python tools/sanitize_target.py auto > gpurun_out/sanitizer/initcheck_auto.log 2>&1
echo "initcheck auto $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer/initcheck_auto.log | tail -1)"
timeout 600 $S --tool memcheck --print-limit 20 python tools/sanitize_target.py auto > gpurun_out/sanitizer/memcheck_auto.log 2>&1
echo "memcheck auto $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer/memcheck_auto.log | tail -1)"
