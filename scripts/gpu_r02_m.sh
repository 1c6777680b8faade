#!/bin/bash
# round 2, call M: single-launch candidate scoring for L2-resident graphs (pairs_small.cuh): parity tests + timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py tests/test_gpu_configs.py tests/test_gpu_replay.py tests/test_golden.py tests/test_mex_gateway.py tests/test_reference_goldens.py -m gpu -q --timeout=1200 -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02m_pytest.log; tail -4 gpurun_out/r02m_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02m_pytest.log | cut -c1-300 | head -30
timeout 900 python scripts/time_small.py grid_England transport_Rome oregon_A8 > gpurun_out/r02m_time_small.jsonl 2> gpurun_out/r02m_time_small.err; echo "rc $?"; cat gpurun_out/r02m_time_small.jsonl; tail -5 gpurun_out/r02m_time_small.err
timeout 900 python scripts/replay_unweighted.py 2>&1 | tail -4 | cut -c1-700 | tee gpurun_out/r02m_replay_unweighted.txt
