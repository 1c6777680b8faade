#!/bin/bash
# round 2, call I: in-place edge edits (delta entries), MEX gateway extensions, full GPU test tier
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest.log; tail -4 gpurun_out/r02i_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02i_pytest.log | head -30
python scripts/time_set_edges.py 2>&1 | tail -1 | tee gpurun_out/r02i_time_set_edges.json
KR_SET_EDGES_REBUILD=1 python scripts/time_set_edges.py 2>&1 | tail -1 | tee gpurun_out/r02i_time_set_edges_rebuild.json
python scripts/replay_unweighted.py 2>&1 | tail -3 | cut -c1-600
