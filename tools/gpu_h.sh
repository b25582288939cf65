#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_o.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_o.log; tail -40 gpurun_out/pytest_o.log | cut -c1-1200
python tools/ncu_target.py cfg4 auto 8 | tail -1
python tools/ncu_target.py cfg2 auto 8 | tail -1
SKYOPT_EXP=8 python tools/ncu_target.py cfg4 auto 8 | tail -1
SKYOPT_TRACE=gpurun_out/trace_cfg4.txt python tools/ncu_target.py cfg4 auto 6 | tail -1
SKYOPT_TRACE=gpurun_out/trace_cfg2.txt python tools/ncu_target.py cfg2 auto 6 | tail -1
python tools/trace2.py gpurun_out/trace_cfg4.txt > gpurun_out/trace_cfg4_summary.txt 2>&1; tail -30 gpurun_out/trace_cfg4_summary.txt
