#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_replay.py -m gpu -q --timeout=900 > gpurun_out/pytest_replay.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_replay.log; tail -5 gpurun_out/pytest_replay.log; grep "^E " gpurun_out/pytest_replay.log | head
