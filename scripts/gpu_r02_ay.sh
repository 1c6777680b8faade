#!/bin/bash
# round 2, call AY: ncu evidence of the FINAL build - full capture + launch list of the default SpMM inside the bench (m = 2 steps),
# launch list of the exact candidate path (C5 shape, 4096 candidates), full capture of the fused Taylor-term kernel (C4 shape, scale 22)
mkdir -p gpurun_out
export KR_BENCH_M=2 KR_BENCH_EDGES=0 KR_BENCH_C4=0
python bench.py --steps 1 --warmup 3 > gpurun_out/r02ay_bench_m2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 6 -c 1 -o gpurun_out/r02ay_prof_spmm \
    python bench.py --steps 1 --warmup 3 > gpurun_out/r02ay_ncu_spmm.log 2>&1
echo "spmm capture rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ay_launches_bench_k512_m2.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/r02ay_ncu_launch.log 2>&1
echo "launch list rc $?"
unset KR_BENCH_M KR_BENCH_EDGES KR_BENCH_C4
python scripts/bench_edges.py --ncand 4096 > gpurun_out/r02ay_edges_4096cand.json 2> gpurun_out/r02ay_edges.err && cat gpurun_out/r02ay_edges_4096cand.json &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02ay_launches_edges_4096cand.csv \
    python scripts/bench_edges.py --ncand 4096 > gpurun_out/r02ay_ncu_edges.log 2>&1
echo "edges launch list rc $?"
python scripts/bench_expmv.py --scale 22 --device-gen > gpurun_out/r02ay_expmv_rmat22.json 2> gpurun_out/r02ay_expmv.err && tail -1 gpurun_out/r02ay_expmv_rmat22.json | cut -c1-600 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 60 -c 1 -o gpurun_out/r02ay_prof_taylor \
    python scripts/bench_expmv.py --scale 22 --device-gen > gpurun_out/r02ay_ncu_taylor.log 2>&1
echo "taylor capture rc $?"
ls -la gpurun_out | grep r02ay
