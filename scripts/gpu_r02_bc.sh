#!/bin/bash
# round 2, call BC: ncu (source-level) of entries_local_kernel with the QL solve
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:entries_local_kernel -s 1 -c 1 -o gpurun_out/r02bc_prof_entries_local_ql \
    python scripts/bench_c2.py --check 0 > gpurun_out/r02bc_ncu.log 2>&1; echo "ncu rc $?"; tail -2 gpurun_out/r02bc_ncu.log | cut -c1-300
