#!/bin/bash
# ncu --set full captures of the other bandwidth kernels of the path (one launch each)
mkdir -p gpurun_out
export KR_BENCH_M=2 KR_BENCH_EDGES=0
ncu --set full --clock-control none --import-source on -k regex:combine3_norm_kernel -s 4 -c 1 -o gpurun_out/prof_combine3 \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_combine3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pair_update_kernel|spmm_kernel" -s 6 -c 3 -o gpurun_out/prof_pairs \
    python scripts/bench_edges.py --ncand 1024 > gpurun_out/ncu_pairs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 60 -c 1 -o gpurun_out/prof_taylor \
    python scripts/bench_expmv.py --scale 22 --nnz 67108864 --q 64 > gpurun_out/ncu_taylor.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_combine3.log gpurun_out/ncu_pairs.log gpurun_out/ncu_taylor.log
