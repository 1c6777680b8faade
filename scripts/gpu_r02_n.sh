#!/bin/bash
mkdir -p gpurun_out
python scripts/time_pairs_small.py oregon_A8 transport_Rome oregon_A7 > gpurun_out/r02n_time_pairs_small_multisect_ilp.jsonl 2>&1; cat gpurun_out/r02n_time_pairs_small_multisect_ilp.jsonl
python -m pytest tests/test_gpu_krylov.py tests/test_gpu_configs.py tests/test_gpu_replay.py tests/test_golden.py tests/test_mex_gateway.py tests/test_reference_goldens.py -m gpu -q --timeout=1200 > gpurun_out/r02n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_pytest.log; tail -4 gpurun_out/r02n_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02n_pytest.log | cut -c1-300 | head -30
