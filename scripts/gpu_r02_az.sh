#!/bin/bash
# round 2, call AZ: the local (one CTA per space) path of function_multiple_entries - parity tests, then config C2 at full size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_krylov.py tests/test_reference_goldens.py tests/test_gpu_differential.py tests/test_gpu_configs.py -m gpu -q --timeout=300 -k "entries or gradient or grad or hessian or golden or vermont or Vermont" > gpurun_out/r02az_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02az_pytest.log; tail -3 gpurun_out/r02az_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02az_pytest.log | cut -c1-300 | head -20
timeout 300 python scripts/bench_c2.py --check 24 > gpurun_out/r02az_c2_local.json 2> gpurun_out/r02az_c2_local.err; echo "c2 local rc $?"; cat gpurun_out/r02az_c2_local.json; tail -3 gpurun_out/r02az_c2_local.err
