#!/bin/bash
# prefetch build: GPU test tier + default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_k.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k.log; tail -4 gpurun_out/pytest_k.log
python bench.py > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo; tail -c 3000 gpurun_out/bench_k.json
