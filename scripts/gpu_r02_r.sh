#!/bin/bash
# round 2, call R: CholQR2 + Householder sign reconstruction in the wide-block path (thin_qr): full GPU tier, A/B timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02r_pytest.log; tail -3 gpurun_out/r02r_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02r_pytest.log | cut -c1-300 | head -20
timeout 900 python scripts/time_small.py grid_England transport_Rome oregon_A8 > gpurun_out/r02r_time_small_cholqr.jsonl 2> gpurun_out/r02r_time_small.err; cat gpurun_out/r02r_time_small_cholqr.jsonl | cut -c1-420; tail -3 gpurun_out/r02r_time_small.err
python scripts/gram_case.py 2>&1 | tail -1
KR_QR_HOUSEHOLDER=1 python scripts/gram_case.py 2>&1 | tail -1
