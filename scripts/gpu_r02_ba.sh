#!/bin/bash
# round 2, call BA: config C2 with the time breakdown (local path against the dense batch in one process), ncu of entries_local_kernel
mkdir -p gpurun_out
timeout 300 python scripts/bench_c2.py --check 12 > gpurun_out/r02ba_c2_breakdown.json 2> gpurun_out/r02ba_c2.err; echo "c2 rc $?"; cat gpurun_out/r02ba_c2_breakdown.json; tail -3 gpurun_out/r02ba_c2.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:entries_local_kernel -s 1 -c 1 -o gpurun_out/r02ba_prof_entries_local \
    python scripts/bench_c2.py --check 0 > gpurun_out/r02ba_ncu.log 2>&1; echo "ncu rc $?"
