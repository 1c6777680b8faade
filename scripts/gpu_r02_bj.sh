#!/bin/bash
# round 2, call BJ: ncu of the final entries_local_kernel (tier-1 launch of the C2 call: 83 184 spaces, five CTAs per SM)
mkdir -p gpurun_out
KR_C2_SKIP_DENSE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:entries_local_kernel -s 1 -c 1 -o gpurun_out/r02bj_prof_entries_local_final \
    python scripts/bench_c2.py --check 0 > gpurun_out/r02bj_ncu.log 2>&1; echo "ncu rc $?"; tail -1 gpurun_out/r02bj_ncu.log | cut -c1-200
