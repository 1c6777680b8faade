#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_screen.py -m gpu -q --timeout=900 -x > gpurun_out/r02u_pytest.log 2>&1; tail -25 gpurun_out/r02u_pytest.log | cut -c1-400
