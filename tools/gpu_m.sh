#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_target.py cfg4 auto 8 | tail -1
python tools/ncu_target.py cfg2 auto 8 | tail -1
SKYOPT_TRACE=gpurun_out/trace_cfg4.txt python tools/ncu_target.py cfg4 auto 6 | tail -1
python tools/trace2.py gpurun_out/trace_cfg4.txt | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_big_configs.py tests/test_gpu_catalog_api.py tests/test_gpu_random_dag.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_target.py auto 2>&1 | grep -E "RACECHECK SUMMARY|mismatches"
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_target.py auto 2>&1 | grep -E "ERROR SUMMARY|mismatches"
