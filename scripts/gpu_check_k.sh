#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py -m gpu -q --timeout=900 -k "trace_fun_update" > gpurun_out/pytest_k3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k3.log; tail -6 gpurun_out/pytest_k3.log; grep "^E " gpurun_out/pytest_k3.log | head
