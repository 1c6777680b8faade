#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_b.log
tail -40 gpurun_out/pytest_b.log
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_b_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_b_all.log
tail -15 gpurun_out/pytest_b_all.log
