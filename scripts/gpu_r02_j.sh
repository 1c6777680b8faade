#!/bin/bash
# round 2, call J: hunt for reads of uninitialised pool memory (KR_POOL_POISON) + new config tests
mkdir -p gpurun_out
KR_POOL_POISON=1 python -m pytest tests/test_gpu_krylov.py tests/test_mex_gateway.py tests/test_gpu_expmv.py tests/test_gpu_spmm.py -m gpu -q --timeout=1200 > gpurun_out/r02j_pytest_poison.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest_poison.log; tail -4 gpurun_out/r02j_pytest_poison.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02j_pytest_poison.log | cut -c1-300 | head -40
python -m pytest tests/test_gpu_configs.py -m gpu -q --timeout=1800 > gpurun_out/r02j_pytest_configs.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest_configs.log; tail -4 gpurun_out/r02j_pytest_configs.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02j_pytest_configs.log | cut -c1-300 | head -20
