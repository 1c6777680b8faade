#!/bin/bash
# round 2, call K (re-entry baseline): full GPU test tier, smoke, small-graph timings, launch list of the bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest.log; tail -4 gpurun_out/r02k_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02k_pytest.log | cut -c1-300 | head -30
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python scripts/time_small.py grid_England transport_Rome oregon_A8 > gpurun_out/r02k_time_small.jsonl 2> gpurun_out/r02k_time_small.err; echo "rc $?"; cat gpurun_out/r02k_time_small.jsonl; tail -5 gpurun_out/r02k_time_small.err
