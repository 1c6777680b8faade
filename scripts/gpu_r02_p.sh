#!/bin/bash
# round 2, call P (2 GPUs): in-process replicas behind the C ABI + torchrun bench at N = 2 (weak + strong + C5 split)
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_replicas.py tests/test_gpu_krylov.py -m gpu -q --timeout=900 > gpurun_out/r02p_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02p_pytest_2gpu.log; tail -4 gpurun_out/r02p_pytest_2gpu.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02p_pytest_2gpu.log | cut -c1-300 | head
python scripts/bench_replicas.py > gpurun_out/r02p_bench_replicas.json 2> gpurun_out/r02p_bench_replicas.err; echo "replicas rc $?"; cat gpurun_out/r02p_bench_replicas.json; tail -3 gpurun_out/r02p_bench_replicas.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02p_bench_2gpu.json 2> gpurun_out/r02p_bench_2gpu.err; echo "bench2 rc $?"; tail -3 gpurun_out/r02p_bench_2gpu.err; cut -c1-600 gpurun_out/r02p_bench_2gpu.json
