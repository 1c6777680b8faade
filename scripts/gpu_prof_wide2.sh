#!/bin/bash
for jm in 32 16 64; do echo "== KR_JACOBI_MAX=$jm"; KR_JACOBI_MAX=$jm ./scripts/gpu_prof_wide.sh 2>&1 | grep "total ms\|====" ; done
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_k4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k4.log; tail -5 gpurun_out/pytest_k4.log; grep "^E " gpurun_out/pytest_k4.log | head
