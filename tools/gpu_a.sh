#!/bin/bash
# first light of the round-2 kernels: parity suite, then a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not cfg5" > gpurun_out/pytest_a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_a.log
tail -15 gpurun_out/pytest_a.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_a.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/bench_a.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','phases_ms','optimize_p50_ms','optimize_cold_p50_ms')}, d['roofline'])
    print('cfg2', {k:d['latency_cfg2'][k] for k in ('ms_per_step','phases_ms','optimize_p50_ms','optimize_cold_p50_ms')})
except Exception as e: print('no bench line', e)
P
