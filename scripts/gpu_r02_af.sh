#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py tests/test_gpu_configs.py tests/test_gpu_expmv.py tests/test_golden.py -m gpu -q --timeout=900 -k "normest or centrality or c5 or c1 or golden or miobi" 2>&1 | tail -4 | cut -c1-300
timeout 900 python scripts/time_small.py grid_England transport_Rome oregon_A8 > gpurun_out/r02af_time_small.jsonl 2> gpurun_out/r02af_time_small.err; cat gpurun_out/r02af_time_small.jsonl | cut -c1-330; tail -3 gpurun_out/r02af_time_small.err
