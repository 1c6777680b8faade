#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py tests/test_gpu_replay.py -m gpu -q --timeout=900 > gpurun_out/pytest_k7.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k7.log; tail -4 gpurun_out/pytest_k7.log; grep "^E " gpurun_out/pytest_k7.log | head
timeout 300 python scripts/time_small.py grid_England transport_Rome oregon_A8 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['graph'], 'centrality b200 %.1f ms oracle %.1f ms' % (d['b200']['centrality_ms'], d['oracle']['centrality_ms']))
"
