#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_golden.py -m gpu -q > gpurun_out/pytest_w.log 2>&1; echo "exit $?"; tail -8 gpurun_out/pytest_w.log
