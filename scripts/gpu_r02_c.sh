#!/bin/bash
# round 2, call C: streaming candidate pipeline (pairs.cuh) - parity tests + C5-shape timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py tests/test_gpu_replay.py tests/test_gpu_spmm.py -m gpu -q --timeout=900 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log; tail -5 gpurun_out/r02c_pytest.log; grep "^E " gpurun_out/r02c_pytest.log | head -20
timeout 600 python scripts/bench_edges.py --ncand 4096 2>&1 | tail -1 | tee gpurun_out/r02c_edges_4096.json
KR_B200_LIB=$PWD/krylov_robustness_b200/libkrylov_r01.so timeout 600 python scripts/bench_edges.py --ncand 4096 2>&1 | tail -1 | tee gpurun_out/r02c_edges_4096_r01.json
