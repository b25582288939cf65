#!/bin/bash
# bench.py under torchrun on N GPUs of one box (gpurun --gpus N).
N=${1:-2}
mkdir -p gpurun_out/multi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi/bench_n$N.json 2> gpurun_out/multi/bench_n$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/multi/bench_n$N.json; tail -3 gpurun_out/multi/bench_n$N.err
