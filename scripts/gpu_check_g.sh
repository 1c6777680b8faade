#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py -m gpu -q -x > gpurun_out/pytest_s.log 2>&1; tail -3 gpurun_out/pytest_s.log
python scripts/replay_unweighted.py --graphs oregon_A0,oregon_A8,transport_Rome 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['graph'], d['method'], round(d['time_s'],4), round(d.get('edges_per_s',0)))
"
