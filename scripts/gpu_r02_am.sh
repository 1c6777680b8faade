#!/bin/bash
# round 2, call AM: the generator script of the reference goldens run through the drop-in MATLAB wrappers + MEX gateway
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dropin_matlab.py tests/test_reference_goldens.py -m gpu -q --timeout=900 > gpurun_out/r02am_pytest_dropin.log 2>&1; echo "dropin exit $?"; tail -40 gpurun_out/r02am_pytest_dropin.log | cut -c1-300
