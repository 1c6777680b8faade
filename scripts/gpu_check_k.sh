#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_k5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k5.log; tail -4 gpurun_out/pytest_k5.log; grep "^E " gpurun_out/pytest_k5.log | head
./scripts/gpu_prof_wide.sh 2>&1 | grep "total ms\|====\|kr wide" | tail -6
