#!/bin/bash
# round 2, call BB: local path of function_multiple_entries with the warp-level tridiagonal QL solve - parity tests, C2 A/B against the Jacobi solve
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_krylov.py tests/test_reference_goldens.py tests/test_gpu_differential.py tests/test_gpu_configs.py -m gpu -q --timeout=300 -k "entries or gradient or grad or hessian or golden or vermont or Vermont" > gpurun_out/r02bb_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02bb_pytest.log; tail -3 gpurun_out/r02bb_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02bb_pytest.log | cut -c1-300 | head -20
timeout 300 python scripts/bench_c2.py --check 24 > gpurun_out/r02bb_c2_ql.json 2> gpurun_out/r02bb_c2_ql.err; echo "c2 ql rc $?"; cat gpurun_out/r02bb_c2_ql.json; tail -3 gpurun_out/r02bb_c2_ql.err
KR_ENTRIES_LOCAL_JACOBI=1 timeout 300 python scripts/bench_c2.py --check 0 > gpurun_out/r02bb_c2_jacobi.json 2> gpurun_out/r02bb_c2_jacobi.err; echo "c2 jacobi rc $?"; cat gpurun_out/r02bb_c2_jacobi.json
