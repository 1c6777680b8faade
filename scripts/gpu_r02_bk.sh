#!/bin/bash
# round 2, call BK: local path with the scaled-Taylor projected solve - parity tests, C2 A/B against the QL solve
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_krylov.py tests/test_reference_goldens.py tests/test_gpu_differential.py tests/test_gpu_configs.py -m gpu -q --timeout=300 -k "entries or gradient or grad or hessian or golden or vermont or Vermont" > gpurun_out/r02bk_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02bk_pytest.log; tail -3 gpurun_out/r02bk_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02bk_pytest.log | cut -c1-300 | head -20
KR_C2_SKIP_DENSE=1 timeout 300 python scripts/bench_c2.py --check 24 > gpurun_out/r02bk_c2_taylor.json 2> gpurun_out/r02bk_c2.err; echo "c2 rc $?"; cat gpurun_out/r02bk_c2_taylor.json; tail -3 gpurun_out/r02bk_c2.err
KR_ENTRIES_LOCAL_SOLVE=1 KR_C2_SKIP_DENSE=1 timeout 300 python scripts/bench_c2.py --check 0 > gpurun_out/r02bk_c2_ql.json 2>/dev/null; echo "c2 ql rc $?"; cut -c1-700 gpurun_out/r02bk_c2_ql.json
