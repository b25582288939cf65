#!/bin/bash
# parity suite, traces, then the full bench (both arms) as the driver runs it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_h.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_h.log
tail -3 gpurun_out/pytest_h.log
for wl in cfg4 cfg2; do
  SKYOPT_TRACE=gpurun_out/trace_$wl.bin timeout 300 python bench.py --workload $wl --no-extras --steps 10 --warmup 3 > gpurun_out/bench_h_$wl.json 2> gpurun_out/bench_h_$wl.err
  echo "== $wl rc=$?"; tail -c 600 gpurun_out/bench_h_$wl.err
  python - <<P
import json
d=json.loads(open('gpurun_out/bench_h_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','optimize_cold_p50_ms','optimize_warm_p50_ms','phases_ms_separate_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'])
P
  python tools/trace2.py gpurun_out/trace_$wl.bin
done
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "full bench rc=$?"; tail -c 800 gpurun_out/bench_full.err
(time timeout 900 python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref bench rc=$?"; tail -c 400 gpurun_out/bench_ref.err
