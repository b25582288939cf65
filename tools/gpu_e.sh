#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_j.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_j.log; tail -3 gpurun_out/pytest_j.log
for m in cold fresh warm; do timeout 120 python tools/profile_e2e.py cfg2 $m > gpurun_out/prof_cfg2_$m.txt 2>&1; head -3 gpurun_out/prof_cfg2_$m.txt | grep p50; done
timeout 120 python tools/profile_e2e.py cfg4 cold > gpurun_out/prof_cfg4_cold.txt 2>&1; grep p50 gpurun_out/prof_cfg4_cold.txt
# fused-step trace (first timing loop of the process)
for wl in cfg4 cfg2; do
  SKYOPT_TRACE=gpurun_out/trace_$wl.bin timeout 300 python tools/ncu_target.py $wl auto 8 | tail -1
  python tools/trace2.py gpurun_out/trace_$wl.bin
done
# launch lists
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file gpurun_out/launches_fused_cfg4.csv python tools/ncu_target.py cfg4 auto 6 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file gpurun_out/launches_split_cfg4.csv python tools/ncu_target.py cfg4 fast-split 6 > gpurun_out/ncu2.log 2>&1
# full captures of the scan kernel: cfg4 as it runs, and the stress row
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan2_kernel -s 3 -c 2 -o gpurun_out/scan2_cfg4 python tools/ncu_target.py cfg4 fast-split 6 > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan2_kernel -s 2 -c 2 -o gpurun_out/scan2_stress python tools/ncu_target.py stress fast-split-noprune 4 > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/step_cfg4 python tools/ncu_target.py cfg4 auto 6 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu3.log gpurun_out/ncu4.log gpurun_out/ncu5.log
