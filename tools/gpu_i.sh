#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_p.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_p.log; tail -15 gpurun_out/pytest_p.log | cut -c1-600
python tools/ncu_target.py cfg4 auto 8 | tail -1
python tools/ncu_target.py cfg2 auto 8 | tail -1
SKYOPT_TRACE=gpurun_out/trace_cfg4.txt python tools/ncu_target.py cfg4 auto 6 | tail -1
python tools/trace2.py gpurun_out/trace_cfg4.txt | tail -12
python tools/profile_e2e.py cfg2 cold > gpurun_out/prof_cfg2_cold.txt 2>&1; head -30 gpurun_out/prof_cfg2_cold.txt | cut -c1-150
python tools/profile_e2e.py cfg4 cold > gpurun_out/prof_cfg4_cold.txt 2>&1; head -3 gpurun_out/prof_cfg4_cold.txt
python tools/profile_e2e.py cfg2 fresh > gpurun_out/prof_cfg2_fresh.txt 2>&1; head -3 gpurun_out/prof_cfg2_fresh.txt
