#!/bin/bash
mkdir -p gpurun_out
python scripts/diag_edge_set_tol.py 2>&1 | tail -4
KR_QR_HOUSEHOLDER=1 python scripts/diag_edge_set_tol.py 2>&1 | tail -4
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ae_pytest.log; tail -3 gpurun_out/r02ae_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02ae_pytest.log | cut -c1-300 | head -20
