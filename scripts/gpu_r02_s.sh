#!/bin/bash
mkdir -p gpurun_out
for g in grid_England oregon_A8; do
KR_PROFILE_WIDE=1 python scripts/fg_case.py $g 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02s_launches_fg_$g.csv python scripts/fg_case.py $g > /dev/null 2>&1
done
python -m pytest tests/test_gpu_replicas.py -m gpu -q 2>&1 | tail -2
