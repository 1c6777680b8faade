#!/bin/bash
# round 2, call AQ: device vs the reference-run C1 trace_exp goldens, the Misc config tests with the adjacent-twin rule
mkdir -p gpurun_out
python -m pytest tests/test_reference_goldens.py tests/test_gpu_configs.py -m gpu -q --timeout=900 > gpurun_out/r02aq_pytest.log 2>&1; echo "exit $?"; tail -5 gpurun_out/r02aq_pytest.log | cut -c1-300; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02aq_pytest.log | cut -c1-300 | head
