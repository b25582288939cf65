#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_m.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_m.log; tail -40 gpurun_out/pytest_m.log | cut -c1-1500
python tools/ncu_target.py stress fast-split-noprune 8 | tail -1
python tools/ncu_target.py cfg4 fast-split 8 | tail -1
python tools/ncu_target.py cfg4 auto 8 | tail -1
SKYOPT_SCAN2_BLOCKS_PER_SM=3 python tools/ncu_target.py stress fast-split-noprune 8 | tail -1
SKYOPT_SCAN2_BLOCKS_PER_SM=2 python tools/ncu_target.py stress fast-split-noprune 8 | tail -1
