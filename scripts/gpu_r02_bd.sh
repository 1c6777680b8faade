#!/bin/bash
# round 2, call BD: local path with three CTAs per SM (768-node balls, 4096-entry arena) and rsqrt rotations in the QL solve
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_krylov.py tests/test_reference_goldens.py tests/test_gpu_differential.py tests/test_gpu_configs.py -m gpu -q --timeout=300 -k "entries or gradient or grad or hessian or golden or vermont or Vermont" > gpurun_out/r02bd_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02bd_pytest.log; tail -3 gpurun_out/r02bd_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02bd_pytest.log | cut -c1-300 | head -20
KR_C2_SKIP_DENSE=1 timeout 300 python scripts/bench_c2.py --check 24 > gpurun_out/r02bd_c2_ql3.json 2> gpurun_out/r02bd_c2_ql3.err; echo "c2 rc $?"; cat gpurun_out/r02bd_c2_ql3.json; tail -3 gpurun_out/r02bd_c2_ql3.err
