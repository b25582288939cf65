#!/bin/bash
# Round-2 profile set (profiles/round2_*): launch lists of the bench command,
# --set full captures of the scan kernel (cfg4, cfg2, no-prune stress) and of
# the fused step kernel, and the full bench line.
mkdir -p gpurun_out/prof
N=/usr/local/cuda/bin/ncu
# 1. the bench line itself (never under a profiler)
python bench.py > gpurun_out/prof/bench_full.json 2> gpurun_out/prof/bench_full.err
tail -c 600 gpurun_out/prof/bench_full.json
# 2. launch lists of the same command (fused headline + split-mode roofline loop)
$N --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/prof/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/prof/bench_under_ncu.log 2>&1
# 3. --set full of the scan kernel: cfg4, cfg2 and the no-prune stress row
$N --set full --clock-control none --import-source on -k regex:scan2_kernel -s 4 -c 2 \
   -o gpurun_out/prof/scan2_cfg4 -f python tools/ncu_target.py cfg4 fast-split 8 > /dev/null 2>&1
$N --set full --clock-control none --import-source on -k regex:scan2_kernel -s 4 -c 2 \
   -o gpurun_out/prof/scan2_cfg2 -f python tools/ncu_target.py cfg2 fast-split 8 > /dev/null 2>&1
$N --set full --clock-control none --import-source on -k regex:scan2_kernel -s 4 -c 2 \
   -o gpurun_out/prof/scan2_stress -f python tools/ncu_target.py stress fast-split-noprune 8 > /dev/null 2>&1
# 4. the fused step kernel (cooperative: replayed whole)
$N --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 2 \
   -o gpurun_out/prof/step_cfg4 -f python tools/ncu_target.py cfg4 auto 8 > /dev/null 2>&1
$N --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 2 \
   -o gpurun_out/prof/step_cfg2 -f python tools/ncu_target.py cfg2 auto 8 > /dev/null 2>&1
# 5. launch lists of the two modes on their own
$N --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/prof/launches_split_cfg4.csv python tools/ncu_target.py cfg4 fast-split 8 > /dev/null 2>&1
$N --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/prof/launches_fused_cfg4.csv python tools/ncu_target.py cfg4 auto 8 > /dev/null 2>&1
ls -la gpurun_out/prof
