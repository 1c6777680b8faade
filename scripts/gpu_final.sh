#!/bin/bash
# round-end rehearsal: GPU test tier, smoke, both bench arms
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_final.log; tail -4 gpurun_out/pytest_final.log
python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; tail -c 600 gpurun_out/bench_final_reference.json
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo; tail -c 2500 gpurun_out/bench_final.json
