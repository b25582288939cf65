#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_target.py cfg4 auto 10 | tail -1
python tools/ncu_target.py cfg2 auto 10 | tail -1
python tools/ncu_target.py cfg4 fast-split 10 | tail -1
python tools/ncu_target.py stress fast-split-noprune 8 | tail -1
SKYOPT_TRACE=gpurun_out/trace_cfg4.txt python tools/ncu_target.py cfg4 auto 6 | tail -1
python tools/trace2.py gpurun_out/trace_cfg4.txt | head -9
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_big_configs.py tests/test_gpu_zz_late_clouds.py tests/test_gpu_random_dag.py -m gpu -q -x 2>&1 | tail -3
