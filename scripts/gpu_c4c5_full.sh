#!/bin/bash
# full-size C4 (R-MAT 2^24 / 2^28, 64 RHS) and C5 (10^5 candidate edges on the C3 graph) runs
mkdir -p gpurun_out
timeout 600 python scripts/bench_expmv.py --scale 24 --nnz 268435456 --q 64 --device-gen > gpurun_out/bench_expmv_full.json 2> gpurun_out/bench_expmv_full.err; echo "expmv rc $?"; tail -c 1200 gpurun_out/bench_expmv_full.json; tail -3 gpurun_out/bench_expmv_full.err
timeout 600 python scripts/bench_edges.py --ncand 100000 > gpurun_out/bench_edges_full.json 2> gpurun_out/bench_edges_full.err; echo "edges rc $?"; tail -c 800 gpurun_out/bench_edges_full.json; tail -3 gpurun_out/bench_edges_full.err
