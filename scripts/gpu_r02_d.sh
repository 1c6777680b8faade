#!/bin/bash
# round 2, call D: hand-written wide-block path (no cuBLAS / cuSOLVER): parity tests + small-graph timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log; tail -5 gpurun_out/r02d_pytest.log; grep -E "^E |^FAILED|^ERROR" gpurun_out/r02d_pytest.log | head -40
timeout 600 python scripts/time_small.py grid_England transport_Rome oregon_A8 > gpurun_out/r02d_time_small.jsonl 2>gpurun_out/r02d_time_small.err; tail -3 gpurun_out/r02d_time_small.jsonl | cut -c1-900
