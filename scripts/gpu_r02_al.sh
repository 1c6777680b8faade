#!/bin/bash
# round 2, call AL: the device path against the REFERENCE-RUN goldens (tests/test_reference_goldens.py, new), then the
# whole GPU tier
mkdir -p gpurun_out
python -m pytest tests/test_reference_goldens.py -m gpu -q --timeout=600 > gpurun_out/r02al_pytest_refgold.log 2>&1; echo "refgold exit $?"; tail -25 gpurun_out/r02al_pytest_refgold.log | cut -c1-400
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02al_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02al_pytest.log; tail -3 gpurun_out/r02al_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02al_pytest.log | cut -c1-300 | head -20
