#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_n.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_n.log; tail -60 gpurun_out/pytest_n.log | cut -c1-1200
for D in 2 4 6; do
  for B in 2 3 4; do
    echo "ring depth $D blocks/SM $B"
    SKYOPT_LIBRARY=skypilot_b200/libskyopt_d$D.so SKYOPT_SCAN2_BLOCKS_PER_SM=$B python tools/ncu_target.py stress fast-split-noprune 8 | tail -1
  done
done
for B in 2 3 4; do
  echo "ring depth 3 blocks/SM $B"
  SKYOPT_SCAN2_BLOCKS_PER_SM=$B python tools/ncu_target.py stress fast-split-noprune 8 | tail -1
done
python tools/ncu_target.py cfg4 auto 8 | tail -1
python tools/ncu_target.py cfg2 auto 8 | tail -1
