#!/bin/bash
mkdir -p gpurun_out
python scripts/time_set_edges.py 2>&1 | tail -1
OMP_NUM_THREADS=1 python scripts/time_set_edges.py 2>&1 | tail -1
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_k6.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k6.log; tail -4 gpurun_out/pytest_k6.log; grep "^E " gpurun_out/pytest_k6.log | head
