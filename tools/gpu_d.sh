#!/bin/bash
mkdir -p gpurun_out
for exp in 0 2 4 6; do
for wl in cfg4 cfg2; do
  SKYOPT_EXP=$exp SKYOPT_TRACE=gpurun_out/trace_${wl}_$exp.bin timeout 300 python bench.py --workload $wl --no-extras --steps 20 --warmup 3 > gpurun_out/bench_i_${wl}_$exp.json 2> gpurun_out/bench_i_${wl}_$exp.err
  echo "== exp=$exp $wl rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/bench_i_${wl}_$exp.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step',)}, d['roofline']['kernel_ms'])
P
  python tools/trace2.py gpurun_out/trace_${wl}_$exp.bin | grep -E "end|start|tables"
done; done
