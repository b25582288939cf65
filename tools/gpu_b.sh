#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_g.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_g.log
tail -4 gpurun_out/pytest_g.log
for wl in cfg4 cfg2; do
  SKYOPT_TRACE=gpurun_out/trace_$wl.bin timeout 300 python bench.py --workload $wl --no-latency --steps 10 --warmup 3 --cpu-baseline-steps 1 > gpurun_out/bench_g_$wl.json 2> gpurun_out/bench_g_$wl.err
  echo "== $wl rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/bench_g_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','phases_ms','optimize_p50_ms','optimize_cold_p50_ms')})
P
  python tools/trace2.py gpurun_out/trace_$wl.bin
done
