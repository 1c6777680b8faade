#!/bin/bash
# round 2, call AN: device against the reference-run C1 greedy goldens
mkdir -p gpurun_out
python -m pytest tests/test_reference_goldens.py -m gpu -q --timeout=900 -k c1 > gpurun_out/r02an_pytest_c1.log 2>&1; echo "c1 exit $?"; tail -30 gpurun_out/r02an_pytest_c1.log | cut -c1-400
