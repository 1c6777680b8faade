#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_edges.py > gpurun_out/bench_edges.json 2> gpurun_out/bench_edges.err; cat gpurun_out/bench_edges.json; tail -2 gpurun_out/bench_edges.err
python scripts/bench_edges.py --tolfac 1e-12 > gpurun_out/bench_edges_tight.json 2>> gpurun_out/bench_edges.err; cat gpurun_out/bench_edges_tight.json
python scripts/bench_edges.py --ncand 256 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_edges.csv python scripts/bench_edges.py --ncand 256 > gpurun_out/ncu_edges.log 2>&1
