#!/bin/bash
# round 2, call X: full GPU tier with the best-fit pool, the screened round, the MEX greedy_round op
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02x_pytest.log; tail -3 gpurun_out/r02x_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02x_pytest.log | cut -c1-300 | head -20
python __graft_entry__.py smoke 2>&1 | tail -1
