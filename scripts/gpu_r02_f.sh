#!/bin/bash
# round 2, call F: wide path after norm2 early exit / QR re-tiling / adaptive squarings; per-step parity diagnostic
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py tests/test_gpu_replay.py tests/test_mex_gateway.py tests/test_gpu_expmv.py -m gpu -q --timeout=900 > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest.log; tail -4 gpurun_out/r02f_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02f_pytest.log | head -20
python scripts/diag_wide.py grid_England cosh sinh 2>&1 | tail -16
python scripts/diag_wide.py transport_Rome sinh cosh 2>&1 | tail -16
for w in fun_and_grad tfu_rank2 tfu_set centrality normest edges250; do KR_PROFILE_WIDE=1 python scripts/one_call.py $w grid_England 2>&1 | tail -3; done
KR_PROFILE_WIDE=1 python scripts/one_call.py fun_and_grad transport_Rome 2>&1 | tail -4
