#!/bin/bash
# round 2, call AP (2 GPUs): whole GPU tier on a 2-GPU box (the in-process replica tests run instead of skipping), torchrun
# bench at N = 2 and its reference arm
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02ap_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ap_pytest_2gpu.log; tail -3 gpurun_out/r02ap_pytest_2gpu.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02ap_pytest_2gpu.log | cut -c1-300 | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02ap_bench_2gpu.json 2> gpurun_out/r02ap_bench_2gpu.err; echo "bench2 rc $?"; tail -3 gpurun_out/r02ap_bench_2gpu.err; cut -c1-500 gpurun_out/r02ap_bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02ap_bench_reference_2gpu.json 2> gpurun_out/r02ap_bench_reference_2gpu.err; echo "ref2 rc $?"; cut -c1-300 gpurun_out/r02ap_bench_reference_2gpu.json
