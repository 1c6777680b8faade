#!/bin/bash
# multi-GPU check: the bench under torchrun (what the driver does for N > 1)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_m_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_m_n$N.log
tail -4 gpurun_out/bench_m_n$N.log | cut -c1-1800
free -g | head -2
