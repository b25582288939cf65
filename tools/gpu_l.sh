#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_q.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_q.log; tail -6 gpurun_out/pytest_q.log | cut -c1-400
python tools/profile_e2e.py cfg2 cold > gpurun_out/prof_cfg2_cold.txt 2>&1; head -24 gpurun_out/prof_cfg2_cold.txt | cut -c1-150
python tools/profile_e2e.py cfg4 cold > gpurun_out/prof_cfg4_cold.txt 2>&1; head -2 gpurun_out/prof_cfg4_cold.txt
python tools/profile_e2e.py cfg2 fresh > gpurun_out/prof_cfg2_fresh.txt 2>&1; head -2 gpurun_out/prof_cfg2_fresh.txt
