#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_target.py cfg4 auto 8 | tail -1
python tools/ncu_target.py cfg2 auto 8 | tail -1
SKYOPT_TRACE=gpurun_out/trace_cfg4.txt python tools/ncu_target.py cfg4 auto 6 | tail -1
python tools/trace2.py gpurun_out/trace_cfg4.txt | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_late_clouds.py tests/test_gpu_big_configs.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
bash tools/gpu_sanitize.sh
